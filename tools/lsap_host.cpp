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

// The kernel's decomposition over `lanes` lanes, emulated: every lane-strided piece runs for lane 0 .. lanes-1 in turn, the
// cross-lane reduction is the same xor butterfly over an array of per-lane candidates (all lanes must agree on the winner),
// and the pieces only meet at the points where the kernel has a __syncwarp().
extern "C" int lsap_host_lanes(const float* cost, int ld, int R, int C, int lanes, long long* col4row) {
    using namespace dpft::lsap;
    if (R > kMaxRows || R > C || lanes < 1 || lanes > 32 || (lanes & (lanes - 1))) return -2;
    std::vector<double> v(C), shortest(C);
    std::vector<int> path(C), row4col(C);
    std::vector<unsigned char> sc(C);
    Workspace w;
    w.v = v.data(); w.shortest = shortest.data(); w.path = path.data(); w.row4col = row4col.data(); w.SC = sc.data();
    for (int l = 0; l < lanes; ++l) init_problem(R, C, w, l, lanes);
    int st = 0;
    for (int cur = 0; cur < R && st == 0; ++cur) {
        for (int l = 0; l < lanes; ++l) init_search(R, C, w, l, lanes);
        double min_val = 0.0;
        int i = cur, sink = -1;
        while (sink < 0) {
            w.SR[i] = 1;
            Candidate c[32];
            for (int l = 0; l < lanes; ++l) c[l] = scan_lane(cost, ld, C, i, min_val, w, l, lanes);
            for (int o = lanes / 2; o > 0; o >>= 1) {
                Candidate n[32];
                for (int l = 0; l < lanes; ++l) n[l] = better(c[l ^ o], c[l]) ? c[l ^ o] : c[l];
                for (int l = 0; l < lanes; ++l) c[l] = n[l];
            }
            for (int l = 1; l < lanes; ++l)
                if (c[l].idx != c[0].idx || c[l].val != c[0].val) return -3;       // the lanes disagree
            if (c[0].idx < 0 || c[0].val >= kInf) { st = -1; break; }
            min_val = c[0].val;
            const int j = c[0].idx;
            w.SC[j] = 1;
            if (w.row4col[j] < 0) sink = j;
            else i = w.row4col[j];
        }
        if (st) break;
        for (int l = 0; l < lanes; ++l) update_duals(R, C, cur, min_val, w, l, lanes);
        augment(cur, sink, w);
    }
    for (int i = 0; i < R; ++i) col4row[i] = st == 0 ? w.col4row[i] : -1;
    return st;
}
