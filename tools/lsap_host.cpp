// Host build of dpft_b200/csrc/lsap_core.h (LANES = 1): the algorithm of the device kernel run sequentially, so that it can be
// checked against scipy.optimize.linear_sum_assignment without a GPU (tests/test_lsap.py compiles this file with g++).
#include <vector>

#include "../dpft_b200/csrc/lsap_core.h"

extern "C" int lsap_host(const float* cost, int ld, int R, int C, long long* col4row) {
    if (R > dpft::lsap::kMaxRows || R > C) return -2;
    std::vector<double> v(C), shortest(C);
    std::vector<int> path(C), row4col(C);
    std::vector<unsigned char> sc(C);
    dpft::lsap::Workspace w;
    w.v = v.data(); w.shortest = shortest.data(); w.path = path.data(); w.row4col = row4col.data(); w.SC = sc.data();
    const int st = dpft::lsap::solve<1>(cost, ld, R, C, w, 0);
    for (int i = 0; i < R; ++i) col4row[i] = st == 0 ? w.col4row[i] : -1;
    return st;
}
