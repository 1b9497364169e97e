// bbcp_ipasir.cpp -- libbbcp_ipasir.so: the IPASIR names of the reference's C ABI (ToporIpasir.h:87-232,
// ToporIpasir.cc:118-178) on top of the bbcp_* entry points, so that an IPASIR client which only needs
// propagation under assumptions (a cube filter, a lookahead front end) can be re-linked against the GPU engine
// without source changes. Semantics that differ from a complete solver, all on the safe side:
//   ipasir_solve   20 = unit propagation of the assumptions derives a conflict (UNSAT for sure);
//                  10 = every variable got assigned without conflict (a model: SAT for sure);
//                   0 = fixpoint reached, undecided -- what IPASIR specifies for an interrupted search
//                       (the reference returns it when its stop callback fires, ToporIpasir.cc:44-52);
//   ipasir_val     value at the propagation fixpoint (lit / -lit / 0 = unassigned), also after a 0 result;
//   ipasir_failed  1 for every assumption of the last call that returned 20 (a superset of a core);
//   ipasir_set_terminate / ipasir_set_learn / ipasir_set_parameter  accepted and ignored: no search runs.
#include <cstdint>
#include <cstdlib>
#include <vector>

#include "../../include/bbcp.h"

namespace {
struct Wrapper {
    bbcp_handle *h = nullptr;
    std::vector<int32_t> assumps, last;
    int last_result = 0;
};
}  // namespace

extern "C" {

const char *ipasir_signature() { return "bbcp-b200 (batched BCP engine behind the Topor/IPASIR surface)"; }

void *ipasir_init()
{
    Wrapper *w = new Wrapper();
    w->h = bbcp_create(0);
    return w;
}

void ipasir_release(void *s)
{
    Wrapper *w = static_cast<Wrapper *>(s);
    if (!w) return;
    bbcp_release(w->h);
    delete w;
}

void ipasir_add(void *s, int32_t lit_or_zero) { bbcp_add(static_cast<Wrapper *>(s)->h, lit_or_zero); }

void ipasir_assume(void *s, int32_t lit) { static_cast<Wrapper *>(s)->assumps.push_back(lit); }

int ipasir_solve(void *s)
{
    Wrapper *w = static_cast<Wrapper *>(s);
    for (int32_t l : w->assumps) bbcp_assume(w->h, l);
    w->last.swap(w->assumps);
    w->assumps.clear();
    w->last_result = bbcp_solve(w->h);
    return w->last_result;
}

int32_t ipasir_val(void *s, int32_t lit)
{
    const int v = bbcp_val(static_cast<Wrapper *>(s)->h, lit);
    return v == 1 ? lit : v == -1 ? -lit : 0;
}

int ipasir_failed(void *s, int32_t lit)
{
    Wrapper *w = static_cast<Wrapper *>(s);
    if (w->last_result != 20) return 0;
    for (int32_t l : w->last) if (l == lit) return 1;
    return 0;
}

void ipasir_set_terminate(void *, void *, int (*)(void *)) {}
void ipasir_set_learn(void *, void *, int, void (*)(void *, int32_t *)) {}
void ipasir_set_parameter(void *, const char *, const char *) {}

}  // extern "C"
