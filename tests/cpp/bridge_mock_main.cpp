// Compiles include/bbcp_bridge.hpp against a stand-in with the CTopor surface the bridge uses (no reference
// needed): checks that the template instantiates and links. Running it needs a GPU only past Finalize().
#include <functional>
#include <span>
#include <vector>

#include "bbcp_bridge.hpp"

struct MockSolver {
    std::vector<std::vector<int32_t>> clauses;
    std::function<int(unsigned, bool)> next;
    void AddClause(std::span<int32_t> c) { clauses.emplace_back(c.begin(), c.end()); }
    int Solve(std::span<int32_t>) { int n = 0; if (next) while (next(0, n == 0)) ++n; return n; }
    void SetParallelData(unsigned, std::function<void(unsigned, int)>, std::function<int(unsigned, bool)> getNext) { next = getNext; }
};

int main()
{
    bbcp::CToporWithGpuProbing<MockSolver> b;
    b.AddClause({1, -2});
    b.AddClause({2, 3});
    const long fed = b.GpuProbe();          // -2 without a device: the engine has no CPU path
    if (fed == -2) return 3;
    return b.Solve({}, false) >= 0 ? 0 : 1;
}
