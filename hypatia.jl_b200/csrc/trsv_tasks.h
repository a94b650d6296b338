// Ticket list of the segmented triangular solve (trsv_seg_kernel, chol_kernels.cuh).  Plain C++: shared by chol.cu and
// the CPU-tier emulation (tests/emu/).
//
// dpotrs call site: ldiv!(x, fact, rhs) (qrchol.jl:68).  Forward sweep (trans, U' y = b): block k needs the blocks
// j < k; backward sweep (U x = y): block k needs the blocks j > k.  pos(j) = position of block j in solve order.
// A ticket = (block column k, a run of at most `seg` tiles in solve order); the run that ends at pos(k) - 1 is the
// block's FINAL ticket.  Tickets are sorted by (pos of the last tile they need, pos(k)): the order in which their last
// input becomes available - every ticket's inputs (block flags, partial sums of the same column) belong to earlier
// tickets.
#pragma once
#include <algorithm>
#include <vector>

namespace hypdev {

inline std::vector<TrsvTask> trsv_build_tasks(int nblk, bool trans, int seg, int* maxseg_out) {
    struct Key {
        int ready, posk;
        TrsvTask t;
    };
    std::vector<Key> keys;
    int maxseg = 1;
    for (int posk = 0; posk < nblk; posk++) {
        const int k = trans ? posk : nblk - 1 - posk;
        const int ndep = posk;                               // tiles in solve order: positions 0 .. posk - 1
        const int nseg = std::max(1, (ndep + seg - 1) / seg);
        maxseg = std::max(maxseg, nseg);
        for (int s = 0; s < nseg; s++) {
            const int p0 = s * seg, p1 = std::min(ndep, p0 + seg);      // positions [p0, p1)
            const int nj = std::max(0, p1 - p0);
            const int j0 = trans ? p0 : nblk - 1 - p0;
            const bool fin = s == nseg - 1;
            keys.push_back({nj > 0 ? p1 - 1 : -1, posk, {k, j0, nj, (nseg << 16) | (s << 1) | (fin ? 1 : 0)}});
        }
    }
    std::stable_sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
        return a.ready != b.ready ? a.ready < b.ready : a.posk < b.posk;
    });
    std::vector<TrsvTask> out;
    out.reserve(keys.size());
    for (auto& e : keys) out.push_back(e.t);
    if (maxseg_out) *maxseg_out = maxseg;
    return out;
}

}  // namespace hypdev
