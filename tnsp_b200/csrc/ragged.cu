// Sector-compact lock-step tensors: device-side planning and data movement (tnsp_b200/TAT/ragged.py).
//
// A lock-step batch of Monte-Carlo chains holds block-symmetric tensors whose sector structure differs from chain to chain
// (sampled physical charges, per-chain greedy cuts).  The reference plans every edge operation on the host per tensor
// (TAT/include/TAT/implement/edge_operator.hpp:34-692, contract.hpp:306-620); here the same integer rules run ON THE DEVICE,
// one CTA / warp per chain, from per-index charge labels, so that the host never waits for a size:
//
//   rt_sort_kernel    merged edge of a group of edges: merged indices sorted by summed charge (edge_operator.hpp:321-404)
//   rt_match_kernel   pairing of row / column sectors with row charge + column charge = target (core.hpp:162-190,
//                     contract.hpp:539-560); sector matrices are stored back to back at even offsets
//   rt_repack_kernel  regroup / transpose between two groupings, or from / to the dense index space
//                     (edge_operator.hpp:651-688, multidimension_span.hpp:250-383)
//   rt_gemm_kernel    ONE grouped GEMM over (chain x sector x tile) on the FP64 tensor pipe (contract.hpp:582-616)
//   elementwise       scale / binary / norms over the stored sectors only (scalar.hpp:46-118, tensor.hpp:631-660)
//
// Bounds: sort / match / repack / elementwise are HBM (and integer-issue) bound: algorithmic bytes = 16 B per stored element;
// the GEMM is HBM-bound for cfg2's sectors (k, n <= 31) and tensor-pipe bound from min(m, n, k) ~ 128.
#include "common.cuh"
#include "ragged.cuh"
#include <climits>
#include <cstdint>
#include <cstdlib>

namespace tnsp {

static unsigned long long* g_stats_dev = nullptr;
static int g_stats_on = 0;
static void stats_alloc() {
    if (!g_stats_dev && cudaMalloc(&g_stats_dev, 16 * sizeof(unsigned long long)) == cudaSuccess) cudaMemset(g_stats_dev, 0, 16 * sizeof(unsigned long long));
}
unsigned long long* rt_stats_ptr() { return g_stats_on ? g_stats_dev : nullptr; }
unsigned long long* rt_overflow_ptr() { stats_alloc(); return g_stats_dev ? g_stats_dev + 15 : nullptr; }

// ------------------------------------------------------------------------------------------------
// rt_sort: one CTA per chain
// ------------------------------------------------------------------------------------------------
struct RtEdges {
    const int* lab[8];
    long long lstride[8];
    int dim[8];
    int sign[8];
    int n;
};
constexpr int kHash = 256;

__device__ __forceinline__ int rt_hash(int key) { return (int)(((unsigned)key * 2654435761u) >> 24) & (kHash - 1); }

// (a 1024-thread instantiation for huge groups -- cfg4's 10^6 merged indices on 37 chains leave most SMs idle -- was tried in round 2 and
// withdrawn: its parity case failed on the B200 and there was no GPU time left to find out why)
template <int kSortThreads>
__global__ void __launch_bounds__(kSortThreads) rt_sort_kernel(RtEdges E, int M, int* __restrict__ table, long long tstride) {
    constexpr int kSortWarps = kSortThreads / 32;
    __shared__ int hkey[kHash];
    __shared__ int hval[kHash];
    __shared__ int skeys[RT_SMAX + 1];
    __shared__ int wcnt[kSortWarps][RT_SMAX + 1];
    __shared__ int base[RT_SMAX + 2];
    __shared__ int nsec_s, overflow_s;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* T = table + (long long)b * tstride;
    int* perm = T + RT_HDR;
    int* inv = T + RT_HDR + M;     // first used as scratch for the keys
    for (int h = tid; h < kHash; h += kSortThreads) hkey[h] = RT_EMPTY;
    for (int i = tid; i < kSortWarps * (RT_SMAX + 1); i += kSortThreads) (&wcnt[0][0])[i] = 0;
    if (tid == 0) overflow_s = 0;
    __syncthreads();
    // pass A: keys + set of distinct keys
    for (int r = tid; r < M; r += kSortThreads) {
        int rest = r, key = 0;
        bool dead = false;
        for (int e = E.n - 1; e >= 0; --e) {
            const int d = E.dim[e];
            const int i = rest % d;
            rest /= d;
            const int l = __ldg(E.lab[e] + (long long)b * E.lstride[e] + i);
            if (l >= RT_DEAD_MIN || l <= -RT_DEAD_MIN) dead = true;
            key += E.sign[e] * l;
        }
        if (dead) key = RT_EMPTY;
        inv[r] = key;
        if (!dead) {
            // one lane per distinct key of the warp inserts it (a group has ~10 distinct charges: without the vote every element
            // would hammer the same few shared-memory words with atomics)
            const unsigned peers = __match_any_sync(__activemask(), key);
            if ((int)(__ffs(peers) - 1) == lane) {
                int h = rt_hash(key);
                for (int probe = 0; probe < kHash; ++probe) {
                    if (hkey[h] == key) break;
                    const int old = atomicCAS(&hkey[h], RT_EMPTY, key);
                    if (old == RT_EMPTY || old == key) break;
                    h = (h + 1) & (kHash - 1);
                    if (probe == kHash - 1) overflow_s = 1;
                }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        for (int h = 0; h < kHash; ++h) {
            const int k = hkey[h];
            if (k == RT_EMPTY) continue;
            if (n >= RT_SMAX) { overflow_s = 1; break; }
            int j = n++;
            while (j > 0 && skeys[j - 1] > k) { skeys[j] = skeys[j - 1]; --j; }
            skeys[j] = k;
        }
        nsec_s = n;
    }
    __syncthreads();
    const int nsec = nsec_s;
    for (int h = tid; h < kHash; h += kSortThreads) {
        const int k = hkey[h];
        int v = nsec;
        if (k != RT_EMPTY)
            for (int i = 0; i < nsec; ++i)
                if (skeys[i] == k) { v = i; break; }
        hval[h] = v;
    }
    __syncthreads();
    auto sector = [&](int key) -> int {
        if (key == RT_EMPTY) return nsec;     // dead indices go behind the valid ones
        int h = rt_hash(key);
        for (int probe = 0; probe < kHash; ++probe) {
            if (hkey[h] == key) return hval[h];
            h = (h + 1) & (kHash - 1);
        }
        return nsec;
    };
    // every warp owns a contiguous range of merged indices: per-warp histogram, prefix over warps, stable placement
    int chunk = (M + kSortWarps - 1) / kSortWarps;
    chunk = (chunk + 31) & ~31;
    const int r0 = warp * chunk, r1 = min(M, r0 + chunk);
    for (int rr = r0; rr < r1; rr += 32) {
        const int r = rr + lane;
        const int sec = r < r1 ? sector(inv[r]) : -1 - lane;
        const unsigned peers = __match_any_sync(0xffffffffu, sec);
        if (r < r1 && (int)(__ffs(peers) - 1) == lane) wcnt[warp][sec] += __popc(peers);      // the warp owns its row of counters
        __syncwarp();
    }
    __syncthreads();
    if (tid <= nsec) {
        int tot = 0;
        for (int w = 0; w < kSortWarps; ++w) { const int t = wcnt[w][tid]; wcnt[w][tid] = tot; tot += t; }
        base[tid + 1] = tot;    // counts for now
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int s = 0; s <= nsec; ++s) { const int c = base[s + 1]; base[s] = acc; acc += c; }
        base[nsec + 1] = acc;
        T[0] = overflow_s ? -1 : nsec;
        T[1] = base[nsec];
        for (int s = 0; s < nsec; ++s) T[2 + s] = skeys[s];
        for (int s = 0; s <= nsec; ++s) T[2 + RT_SMAX + s] = base[s];
    }
    __syncthreads();
    for (int s = lane; s <= nsec; s += 32) wcnt[warp][s] += base[s];
    __syncwarp();
    for (int rr = r0; rr < r1; rr += 32) {
        const int r = rr + lane;
        const bool on = r < r1;
        const int sec = on ? sector(inv[r]) : -1 - lane;      // idle lanes never share a group
        const unsigned mask = __match_any_sync(0xffffffffu, sec);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        int pos = 0;
        if (on) pos = wcnt[warp][sec] + rank;
        __syncwarp();
        if (on && rank == 0) wcnt[warp][sec] += __popc(mask);
        __syncwarp();
        if (on) { perm[pos] = r; inv[r] = pos; }
    }
}

// ------------------------------------------------------------------------------------------------
// rt_sort for SMALL groups (merged dimension <= kSortWarpMax: every bond and most two-edge groups of cfg2 -- 6, 36, 216 indices):
// one WARP per chain, no block barrier, no hash table.  The CTA-wide kernel above pays ~10 us of fixed cost per chain (256-slot hash
// initialisation, a serial scan of the slots, four barriers) for a handful of indices.  Here the distinct charges are extracted in
// ascending order by repeated warp-wide minima (<= 64 of them), then the indices are placed chunk by chunk (stable) with one
// __match_any_sync per 32 indices.  Same table layout, same order as rt_sort_kernel.
// ------------------------------------------------------------------------------------------------
constexpr int kSortWarpMax = 2048;
constexpr int kSortWarpPerCta = 4;

__global__ void __launch_bounds__(32 * kSortWarpPerCta) rt_sort_warp_kernel(RtEdges E, int M, int* __restrict__ table, long long tstride, int nbT) {
    __shared__ int s_keys[kSortWarpPerCta][kSortWarpMax];
    __shared__ int s_skey[kSortWarpPerCta][RT_SMAX + 1];
    __shared__ int s_cnt[kSortWarpPerCta][RT_SMAX + 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kSortWarpPerCta + warp;
    if (b >= nbT) return;
    int* keys = s_keys[warp];
    int* skey = s_skey[warp];
    int* cnt = s_cnt[warp];
    int* T = table + (long long)b * tstride;
    int* perm = T + RT_HDR;
    int* inv = T + RT_HDR + M;
    for (int r = lane; r < M; r += 32) {
        int rest = r, key = 0;
        bool dead = false;
        for (int e = E.n - 1; e >= 0; --e) {
            const int d = E.dim[e];
            const int i = rest % d;
            rest /= d;
            const int l = __ldg(E.lab[e] + (long long)b * E.lstride[e] + i);
            if (l >= RT_DEAD_MIN || l <= -RT_DEAD_MIN) dead = true;
            key += E.sign[e] * l;
        }
        keys[r] = dead ? RT_EMPTY : key;
    }
    __syncwarp();
    // distinct charges in ascending order with their counts
    int nsec = 0, acc = 0, overflow = 0;
    long long last = (long long)INT_MIN;            // valid keys are > RT_EMPTY = INT_MIN
    while (true) {
        int mn = INT_MAX, have = 0;
        for (int r = lane; r < M; r += 32) {
            const int k = keys[r];
            if ((long long)k > last) { if (!have || k < mn) mn = k; have = 1; }
        }
        const unsigned any = __ballot_sync(0xffffffffu, have);
        if (!any) break;
        mn = __reduce_min_sync(0xffffffffu, have ? mn : INT_MAX);
        int c = 0;
        for (int r = lane; r < M; r += 32) c += keys[r] == mn;
        c = __reduce_add_sync(0xffffffffu, c);
        if (nsec >= RT_SMAX) { overflow = 1; break; }
        if (lane == 0) { skey[nsec] = mn; cnt[nsec] = acc; T[2 + nsec] = mn; T[2 + RT_SMAX + nsec] = acc; }
        acc += c;
        ++nsec;
        last = mn;
    }
    if (lane == 0) {
        cnt[nsec] = acc;                      // dead indices go behind the valid ones
        T[0] = overflow ? -1 : nsec;
        T[1] = acc;
        T[2 + RT_SMAX + nsec] = acc;
    }
    __syncwarp();
    for (int r0 = 0; r0 < M; r0 += 32) {
        const int r = r0 + lane;
        const bool on = r < M;
        int sec = -1 - lane;                  // idle lanes never share a group
        if (on) {
            const int k = keys[r];
            sec = nsec;
            if (k != RT_EMPTY) {
                int lo = 0, hi = nsec;            // skey[lo] <= k < skey[hi]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (skey[mid] <= k) lo = mid; else hi = mid; }
                sec = (nsec > 0 && skey[lo] == k) ? lo : nsec;      // (after an overflow the surplus charges are parked with the dead ones)
            }
        }
        const unsigned mask = __match_any_sync(0xffffffffu, sec);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        int pos = 0;
        if (on) pos = cnt[sec] + rank;
        __syncwarp();
        if (on && rank == 0) cnt[sec] += __popc(mask);
        __syncwarp();
        if (on) { perm[pos] = r; inv[r] = pos; }
    }
}

// ------------------------------------------------------------------------------------------------
// rt_match: one warp per chain
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rt_match_warp(const int* __restrict__ rt, long long rts, int rs, const int* __restrict__ ct, long long cts, int cs,
                                              const int* __restrict__ t1, int t1s, int s1, const int* __restrict__ t2, int t2s, int s2,
                                              int* __restrict__ match, int* __restrict__ tsum, long long cap, unsigned long long* flag, int b, int lane) {
    const RtTab R(rt + (long long)b * rts), C(ct + (long long)b * cts);
    int t = 0;
    if (t1) t += s1 * t1[(long long)b * t1s];
    if (t2) t += s2 * t2[(long long)b * t2s];
    if (tsum && lane == 0) tsum[b] = t;
    int* Mrow = match + (long long)b * RT_MSTRIDE;
    const int nr = max(R.nsec(), 0), nc = max(C.nsec(), 0);
    int carry = 0;
    for (int i0 = 0; i0 < nr; i0 += 32) {
        const int i = i0 + lane;
        int j = -1, sz = 0;
        if (i < nr) {
            const int want = t - rs * R.skey(i);
            for (int jj = 0; jj < nc; ++jj)
                if (cs * C.skey(jj) == want) { j = jj; break; }
            if (j >= 0) { sz = R.count(i) * C.count(j); sz += sz & 1; }
        }
        int incl = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (i < nr) { Mrow[2 + i] = carry + incl - sz; Mrow[3 + RT_SMAX + i] = j; }
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    const bool over = cap > 0 && carry > cap;
    if (over) {
        __syncwarp();
        for (int i = lane; i <= nr; i += 32) Mrow[2 + i] = 0;
        for (int i = lane; i < nr; i += 32) Mrow[3 + RT_SMAX + i] = -1;
        if (lane == 0 && flag) atomicAdd(flag, 1ull);
    }
    if (lane == 0) {
        if (!over) Mrow[2 + nr] = carry;
        const bool bad = over || R.nsec() < 0 || C.nsec() < 0;
        Mrow[0] = bad ? 0 : carry;
        Mrow[1] = bad ? 1 : 0;      // sector / capacity overflow flag
    }
}

__global__ void __launch_bounds__(128) rt_match_kernel(const int* __restrict__ rt, long long rts, int rs, const int* __restrict__ ct, long long cts,
                                                       int cs, const int* __restrict__ t1, int t1s, int s1, const int* __restrict__ t2, int t2s,
                                                       int s2, int* __restrict__ match, int* __restrict__ tsum, int nbm, long long cap,
                                                       unsigned long long* flag) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nbm) return;
    rt_match_warp(rt, rts, rs, ct, cts, cs, t1, t1s, s1, t2, t2s, s2, match, tsum, cap, flag, b, lane);
}

// several pairings in one launch: blockIdx.y selects the job
struct RtMatchJobs { tnsp_rt_match_job j[4]; };
__global__ void __launch_bounds__(128) rt_match_multi_kernel(RtMatchJobs jobs, int nbm, unsigned long long* flag) {
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= nbm) return;
    const tnsp_rt_match_job& q = jobs.j[blockIdx.y];
    rt_match_warp(q.rt, q.rt_stride, q.rs, q.ct, q.ct_stride, q.cs, q.t1, q.t1_stride, q.s1, q.t2, q.t2_stride, q.s2, q.match, q.tsum, q.cap, flag, b,
                  lane);
}

// ------------------------------------------------------------------------------------------------
// rt_repack: destination-driven regrouping
// ------------------------------------------------------------------------------------------------
constexpr int kRepackThreads = 256;
constexpr int kPlanEnt = 5;      // per destination edge: dim, source group (0 rows / 1 cols), source stride, magic lo, magic hi
constexpr int kPlanMax = 2 + kPlanEnt * 24;

// exact x / d for x < 2^32 with the precomputed magic m = floor(2^64 / d) + 1 (one 64-bit multiply-high instead of a division)
__device__ __forceinline__ unsigned rt_div(unsigned x, const int* ent) {
    const unsigned long long m = ((unsigned long long)(unsigned)ent[4] << 32) | (unsigned)ent[3];
    return (unsigned)__umul64hi((unsigned long long)x, m);
}
// merged destination index -> contributions to the source's row / column merged indices
// fermionic signs (TAT/ragged_fermi.py): sign(element) = s0 + lin . x + sum_{k<j} Q_kj x_k x_j over the parity bits x of the
// destination's indexed edges; quad[k] = mask of the j > k with Q_kj = 1, per_chain[b] = lin bits | s0 << 31
struct RtSign {
    const int* quad;
    const int* per_chain; long long pcs;
    const int* lab[24]; long long lst[24];
    int fermi;                  // bit i: component i of a packed label is fermionic
};
__device__ __forceinline__ unsigned rt_label_parity(int label, int fermi) {
    const int c0 = ((label + 32768) & 0xffff) - 32768;
    const int c1 = (label - c0) >> 16;
    return (unsigned)(((fermi & 1) ? c0 : 0) ^ ((fermi & 2) ? c1 : 0)) & 1u;
}
// as rt_decode, and the parity bits of the decoded edges (bit k for plan entry k)
__device__ __forceinline__ unsigned rt_decode_bits(const int* sp, int first, int last, unsigned idx, unsigned& sr, unsigned& sc, const RtSign& sg,
                                                   int b) {
    unsigned bits = 0;
    for (int k = last - 1; k >= first; --k) {
        const int* ent = sp + 2 + kPlanEnt * k;
        unsigned r = 0;
        if (ent[0] > 1) {
            const unsigned qd = rt_div(idx, ent);
            r = idx - qd * (unsigned)ent[0];
            idx = qd;
        }
        if (ent[1]) sc += r * (unsigned)ent[2]; else sr += r * (unsigned)ent[2];
        bits |= rt_label_parity(__ldg(sg.lab[k] + (long long)b * sg.lst[k] + r), sg.fermi) << k;
    }
    return bits;
}
__device__ __forceinline__ void rt_decode(const int* sp, int first, int last, unsigned idx, unsigned& sr, unsigned& sc) {
    for (int k = last - 1; k >= first; --k) {
        const int* ent = sp + 2 + kPlanEnt * k;
        if (ent[0] <= 1) continue;
        const unsigned qd = rt_div(idx, ent);
        const unsigned r = idx - qd * (unsigned)ent[0];
        idx = qd;
        if (ent[1]) sc += r * (unsigned)ent[2]; else sr += r * (unsigned)ent[2];
    }
}

struct RtRepackArgs {
    const int* plan; RtForm S, D; RtSpec spec; double* dst; long long dst_stride; long long dense_size;
};
struct RtRepackPair { RtRepackArgs a[2]; };

template <bool SRC_DENSE, bool DST_DENSE>
__device__ __forceinline__ void rt_repack_body(const int* __restrict__ plan, const RtForm& S, const RtForm& D, const RtSpec& spec,
                                               double* __restrict__ dst, long long dst_stride, long long dense_size, unsigned long long* stats) {
    __shared__ int sp[kPlanMax];
    __shared__ int hdr[4][RT_HDR];
    __shared__ int mt[2][RT_MSTRIDE];
    const int b = blockIdx.y, tid = threadIdx.x;
    const int n_ent = plan[0] + plan[1];
    for (int i = tid; i < 2 + kPlanEnt * n_ent; i += kRepackThreads) sp[i] = plan[i];
    if (!SRC_DENSE) {
        for (int i = tid; i < RT_HDR; i += kRepackThreads) { hdr[0][i] = S.rt[b * S.rts + i]; hdr[1][i] = S.ct[b * S.cts + i]; }
        for (int i = tid; i < RT_MSTRIDE; i += kRepackThreads) mt[0][i] = S.match[b * S.mts + i];
    }
    if (!DST_DENSE) {
        for (int i = tid; i < RT_HDR; i += kRepackThreads) { hdr[2][i] = D.rt[b * D.rts + i]; hdr[3][i] = D.ct[b * D.cts + i]; }
        if (!spec.on)
            for (int i = tid; i < RT_MSTRIDE; i += kRepackThreads) mt[1][i] = D.match[b * D.mts + i];
    }
    __syncthreads();
    if (!DST_DENSE && spec.on) rt_match_cta(hdr[2], hdr[3], spec, b, mt[1], blockIdx.x == 0);
    const int nr = sp[0], nc = sp[1];
    const RtTab sR(hdr[0]), sC(hdr[1]), dR(hdr[2]), dC(hdr[3]);
    const RtMatch sM(mt[0]), dM(mt[1]);
    const long long total = DST_DENSE ? dense_size : (long long)dM.size();
    const double* src = S.data + (long long)b * S.dstride;
    double* out = dst + (long long)b * dst_stride;
    if (stats && blockIdx.x == 0 && tid == 0) atomicAdd(&stats[3], (unsigned long long)total);
    const int* s_inv_r = SRC_DENSE ? nullptr : S.rt + b * S.rts + RT_HDR + S.M;
    const int* s_inv_c = SRC_DENSE ? nullptr : S.ct + b * S.cts + RT_HDR + S.N;
    const int* d_perm_r = DST_DENSE ? nullptr : D.rt + b * D.rts + RT_HDR;
    const int* d_perm_c = DST_DENSE ? nullptr : D.ct + b * D.cts + RT_HDR;
    for (long long e = (long long)blockIdx.x * kRepackThreads + tid; e < total; e += (long long)gridDim.x * kRepackThreads) {
        long long rp, cp;
        if (DST_DENSE) {
            rp = e; cp = 0;
        } else {
            // sector of destination element e
            int lo = 0, hi = dR.nsec();
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (dM.moff(mid) <= e) lo = mid; else hi = mid;
            }
            // sectors without a partner have zero extent: step back to the one that holds e
            const int i = lo;
            const int j = dM.mcol(i);
            const int local = (int)(e - dM.moff(i));
            const int n = j >= 0 ? dC.count(j) : 0;
            const int m = dR.count(i);
            if (j < 0 || local >= m * n) { out[e] = 0.0; continue; }     // alignment pad
            const int a = local / n, c = local - a * n;
            rp = d_perm_r[dR.sstart(i) + a];
            cp = d_perm_c[dC.sstart(j) + c];
        }
        unsigned sr = 0, sc = 0;
        rt_decode(sp, 0, nr, (unsigned)rp, sr, sc);
        rt_decode(sp, nr, nr + nc, (unsigned)cp, sr, sc);
        double v = 0.0;
        if (SRC_DENSE) {
            v = __ldg(src + (long long)sr + sc);
        } else {
            const int p = s_inv_r[sr], q = s_inv_c[sc];
            if (p < sR.nvalid() && q < sC.nvalid()) {
                const int i = sR.sector_of(p);
                const int j = sM.mcol(i);
                if (j >= 0 && q >= sC.sstart(j) && q < sC.sstart(j + 1))
                    v = __ldg(src + sM.moff(i) + (long long)(p - sR.sstart(i)) * sC.count(j) + (q - sC.sstart(j)));
            }
        }
        out[e] = v;
    }
}

template <bool SRC_DENSE, bool DST_DENSE>
__global__ void __launch_bounds__(kRepackThreads) rt_repack_kernel(RtRepackArgs p, unsigned long long* stats) {
    rt_repack_body<SRC_DENSE, DST_DENSE>(p.plan, p.S, p.D, p.spec, p.dst, p.dst_stride, p.dense_size, stats);
}
// the two operands of a contraction in one launch: blockIdx.z selects the operand
__global__ void __launch_bounds__(kRepackThreads) rt_repack_pair_kernel(RtRepackPair p, unsigned long long* stats) {
    const RtRepackArgs& q = p.a[blockIdx.z];
    rt_repack_body<false, false>(q.plan, q.S, q.D, q.spec, q.dst, q.dst_stride, q.dense_size, stats);
}

// ------------------------------------------------------------------------------------------------
// Regrouping between two sector-compact layouts, tile by tile: a CTA owns RTR destination rows x RTC destination columns of ONE
// destination sector, decodes every row and every column of the tile once (RTR + RTC index decodes into shared memory instead of
// RTR x RTC) and then moves the elements with two additions, two table look-ups and one sector search each.
// ------------------------------------------------------------------------------------------------
constexpr int kRepackTileThreads = 128;      // default CTA sizes of the tile kernels (see repack_threads)
constexpr int kRepackPairThreads = 64;
constexpr int RTR = 128, RTC = 128;      // largest tile extents; the tile shape is chosen per sector: tc = 2^ceil(log2 n) <= 128, tr = 2048 / tc <= 128
struct RtTileDesc { int m, n, tn, start, rstart, cstart, moff, tr, tc; };

template <bool SIGNED, int NT>
__device__ __forceinline__ void rt_repack_tiles(const int* __restrict__ plan, const RtForm& S, const RtForm& D, const RtSpec& spec,
                                                double* __restrict__ dst, long long dst_stride, unsigned long long* stats, const RtSign* sgp) {
    __shared__ unsigned s_fr[SIGNED ? RTR : 1], s_mr[SIGNED ? RTR : 1], s_fc[SIGNED ? RTC : 1], s_bc[SIGNED ? RTC : 1];
    __shared__ int s_quad[SIGNED ? 24 : 1];
    __shared__ int sp[kPlanMax];
    __shared__ int hS[2][RT_HDR];
    __shared__ int mS[RT_MSTRIDE], mD[RT_MSTRIDE];
    __shared__ RtTileDesc dsc[RT_SMAX];
    __shared__ int n_desc, n_items, cnt[2], tot[2];
    __shared__ unsigned rsr[RTR], rsc[RTR], csr[RTC], csc[RTC];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_ent = plan[0] + plan[1];
    for (int i = tid; i < 2 + kPlanEnt * n_ent; i += NT) sp[i] = plan[i];
    __shared__ int hD[2][RT_HDR];
    const int* gDr = D.rt + b * D.rts;
    const int* gDc = D.ct + b * D.cts;
    for (int i = tid; i < RT_HDR; i += NT) {
        hS[0][i] = S.rt[b * S.rts + i]; hS[1][i] = S.ct[b * S.cts + i];
        hD[0][i] = gDr[i]; hD[1][i] = gDc[i];
    }
    for (int i = tid; i < RT_MSTRIDE; i += NT) mS[i] = S.match[b * S.mts + i];
    if (!spec.on)
        for (int i = tid; i < RT_MSTRIDE; i += NT) mD[i] = D.match[b * D.mts + i];
    __syncthreads();
    if (spec.on) rt_match_cta(hD[0], hD[1], spec, b, mD, blockIdx.x == 0);
    {
        const RtTab dR(hD[0]), dC(hD[1]);
        const RtMatch dM(mD);
        const int nsec = max(dR.nsec(), 0);
        int here = 0;
        RtTileDesc d;
        d.m = d.n = d.tn = d.start = d.rstart = d.cstart = d.moff = 0; d.tr = d.tc = 1;
        if (tid < nsec) {
            const int j = dM.mcol(tid);
            if (j >= 0) {
                d.m = dR.count(tid); d.n = dC.count(j);
                if (d.m > 0 && d.n > 0) {
                    int tc = 1;
                    while (tc < d.n && tc < RTC) tc <<= 1;
                    d.tc = tc;
                    d.tr = min(RTR, 2048 / tc);
                    d.tn = (d.n + d.tc - 1) / d.tc;
                    here = ((d.m + d.tr - 1) / d.tr) * d.tn;
                    d.rstart = dR.sstart(tid); d.cstart = dC.sstart(j); d.moff = dM.moff(tid);
                }
            }
        }
        const unsigned has = __ballot_sync(0xffffffffu, here > 0);
        int incl = here;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (warp < 2 && lane == 31) { cnt[warp] = __popc(has); tot[warp] = incl; }
        __syncthreads();
        if (warp < 2 && here > 0) {
            const int slot = (warp ? cnt[0] : 0) + __popc(has & ((1u << lane) - 1u));
            d.start = (warp ? tot[0] : 0) + incl - here;
            dsc[slot] = d;
        }
        if (tid == 0) { n_desc = cnt[0] + cnt[1]; n_items = tot[0] + tot[1]; }
        __syncthreads();
        if (stats && blockIdx.x == 0 && tid == 0) atomicAdd(&stats[3], (unsigned long long)dM.size());
    }
    const int nr = sp[0], nc = sp[1];
    const RtTab sR(hS[0]), sC(hS[1]);
    const RtMatch sM(mS);
    const double* src = S.data + (long long)b * S.dstride;
    double* out = dst + (long long)b * dst_stride;
    const int* s_inv_r = S.rt + b * S.rts + RT_HDR + S.M;
    const int* s_inv_c = S.ct + b * S.cts + RT_HDR + S.N;
    const int* d_perm_r = gDr + RT_HDR;
    const int* d_perm_c = gDc + RT_HDR;
    const int nd = n_desc, total = n_items;
    const int s_nvr = sR.nvalid(), s_nvc = sC.nvalid();
    unsigned lin = 0, s0 = 0;
    if (SIGNED) {
        const int pc = sgp->per_chain[(long long)b * sgp->pcs];
        lin = (unsigned)pc & 0x7fffffffu; s0 = ((unsigned)pc >> 31) & 1u;
        if (tid < nr + nc) s_quad[tid] = sgp->quad[tid];
        __syncthreads();
    }
    int cur = 0;
    for (int item = blockIdx.x; item < total; item += gridDim.x) {
        while (cur + 1 < nd && dsc[cur + 1].start <= item) ++cur;
        const RtTileDesc e = dsc[cur];
        const int t = item - e.start;
        const int r0 = (t / e.tn) * e.tr, c0 = (t % e.tn) * e.tc;
        const int rows = min(e.tr, e.m - r0), cols = min(e.tc, e.n - c0);
        __syncthreads();        // the previous tile's contributions are no longer read
        for (int x_ = tid; x_ < rows + cols; x_ += NT) {
            if (x_ < rows) {
                const int rr = x_;
                unsigned x = 0, y = 0;
                if (SIGNED) {
                    const unsigned bits = rt_decode_bits(sp, 0, nr, (unsigned)d_perm_r[e.rstart + r0 + rr], x, y, *sgp, b);
                    // f_r = lin . bits + quadratic part inside the row group; m_r = rows of Q selected by the bits (for the cross term)
                    unsigned f = __popc(bits & lin) & 1u, mk = 0;
                    for (int k = 0; k < nr; ++k)
                        if ((bits >> k) & 1u) { f ^= __popc(bits & (unsigned)s_quad[k]) & 1u; mk ^= (unsigned)s_quad[k]; }
                    s_fr[rr] = f; s_mr[rr] = mk;
                } else rt_decode(sp, 0, nr, (unsigned)d_perm_r[e.rstart + r0 + rr], x, y);
                rsr[rr] = x; rsc[rr] = y;
            } else {
                const int cc_ = x_ - rows;
                unsigned x = 0, y = 0;
                if (SIGNED) {
                    const unsigned bits = rt_decode_bits(sp, nr, nr + nc, (unsigned)d_perm_c[e.cstart + c0 + cc_], x, y, *sgp, b);
                    unsigned f = __popc(bits & lin) & 1u;
                    for (int k = nr; k < nr + nc; ++k)
                        if ((bits >> k) & 1u) f ^= __popc(bits & (unsigned)s_quad[k]) & 1u;
                    s_fc[cc_] = f; s_bc[cc_] = bits;
                } else rt_decode(sp, nr, nr + nc, (unsigned)d_perm_c[e.cstart + c0 + cc_], x, y);
                csr[cc_] = x; csc[cc_] = y;
            }
        }
        __syncthreads();
        // (a variant that kept four elements per thread in flight -- index look-ups, sector search, data loads and stores in phases --
        // measured SLOWER on cfg2: 54 instead of 40 registers cost two of six resident CTAs, and with ~1.7 k stored elements per
        // chain and regrouping the CTA count in flight, not the loads per thread, is what hides the latency)
        for (int el = tid; el < rows * cols; el += NT) {
            const int a = el / cols, c = el - a * cols;
            const int p = s_inv_r[rsr[a] + csr[c]], q = s_inv_c[rsc[a] + csc[c]];
            double v = 0.0;
            if (p < s_nvr && q < s_nvc) {
                const int i = sR.sector_of(p);
                const int j = sM.mcol(i);
                if (j >= 0 && q >= sC.sstart(j) && q < sC.sstart(j + 1))
                    v = __ldg(src + sM.moff(i) + (long long)(p - sR.sstart(i)) * sC.count(j) + (q - sC.sstart(j)));
            }
            if (SIGNED) {
                const unsigned sgn = s0 ^ s_fr[a] ^ s_fc[c] ^ (__popc(s_mr[a] & s_bc[c]) & 1u);
                if (sgn) v = -v;
            }
            out[e.moff + (long long)(r0 + a) * e.n + c0 + c] = v;
        }
        if (t == 0 && tid == 0 && ((e.m * e.n) & 1)) out[e.moff + (long long)e.m * e.n] = 0.0;
    }
}

template <int NT>
__global__ void __launch_bounds__(NT) rt_repack_tile_kernel(RtRepackArgs p, unsigned long long* stats) {
    rt_repack_tiles<false, NT>(p.plan, p.S, p.D, p.spec, p.dst, p.dst_stride, stats, nullptr);
}
template <int NT>
__global__ void __launch_bounds__(NT) rt_repack_tile_pair_kernel(RtRepackPair p, unsigned long long* stats) {
    const RtRepackArgs& q = p.a[blockIdx.z];
    rt_repack_tiles<false, NT>(q.plan, q.S, q.D, q.spec, q.dst, q.dst_stride, stats, nullptr);
}
// the same regrouping with the fermionic sign of every element (edge_operator.hpp:497-555, 591, 612 evaluated per element)
template <int NT>
__global__ void __launch_bounds__(NT) rt_repack_signed_kernel(RtRepackArgs p, RtSign sg, unsigned long long* stats) {
    rt_repack_tiles<true, NT>(p.plan, p.S, p.D, p.spec, p.dst, p.dst_stride, stats, &sg);
}

// ------------------------------------------------------------------------------------------------
// rt_dot: contraction of ALL edges of two tensors (the closing contraction of a strip, amplitude x hole, ...): per chain the sum
// over the stored elements of D of D[e] * S[the same multi-index].  Same index walk as the regrouping above, but nothing is
// regrouped or written: no merged group over the whole tensor (whose table would be as large as the tensor), no sorted copy.
// One CTA per chain, fixed summation order (deterministic).  Also writes the 1-element result's pairing table.
// ------------------------------------------------------------------------------------------------
constexpr int kDotThreads = 512;
__global__ void __launch_bounds__(kDotThreads) rt_dot_kernel(const int* __restrict__ plan, RtForm S, RtForm D, const int* __restrict__ t1, int t1st,
                                                             int s1, const int* __restrict__ t2, int t2st, int s2, double* __restrict__ out,
                                                             long long out_stride, int* __restrict__ match, int* __restrict__ tsum,
                                                             unsigned long long* stats) {
    __shared__ int sp[kPlanMax];
    __shared__ int hdr[4][RT_HDR];
    __shared__ int mt[2][RT_MSTRIDE];
    __shared__ double red[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n_ent = plan[0] + plan[1];
    for (int i = tid; i < 2 + kPlanEnt * n_ent; i += kDotThreads) sp[i] = plan[i];
    for (int i = tid; i < RT_HDR; i += kDotThreads) {
        hdr[0][i] = S.rt[b * S.rts + i]; hdr[1][i] = S.ct[b * S.cts + i];
        hdr[2][i] = D.rt[b * D.rts + i]; hdr[3][i] = D.ct[b * D.cts + i];
    }
    for (int i = tid; i < RT_MSTRIDE; i += kDotThreads) { mt[0][i] = S.match[b * S.mts + i]; mt[1][i] = D.match[b * D.mts + i]; }
    __syncthreads();
    const int nr = sp[0], nc = sp[1];
    const RtTab sR(hdr[0]), sC(hdr[1]), dR(hdr[2]), dC(hdr[3]);
    const RtMatch sM(mt[0]), dM(mt[1]);
    const long long total = dM.size();
    const double* src = S.data + (long long)b * S.dstride;
    const double* dat = D.data + (long long)b * D.dstride;
    const int* s_inv_r = S.rt + b * S.rts + RT_HDR + S.M;
    const int* s_inv_c = S.ct + b * S.cts + RT_HDR + S.N;
    const int* d_perm_r = D.rt + b * D.rts + RT_HDR;
    const int* d_perm_c = D.ct + b * D.cts + RT_HDR;
    double acc = 0.0;
    for (long long e = tid; e < total; e += kDotThreads) {
        int lo = 0, hi = dR.nsec();
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (dM.moff(mid) <= e) lo = mid; else hi = mid;
        }
        const int i = lo;
        const int j = dM.mcol(i);
        const int local = (int)(e - dM.moff(i));
        const int n = j >= 0 ? dC.count(j) : 0;
        const int m = dR.count(i);
        if (j < 0 || local >= m * n) continue;
        const int a = local / n, c = local - a * n;
        long long rp = d_perm_r[dR.sstart(i) + a];
        long long cp = d_perm_c[dC.sstart(j) + c];
        unsigned sr = 0, sc = 0;
        rt_decode(sp, 0, nr, (unsigned)rp, sr, sc);
        rt_decode(sp, nr, nr + nc, (unsigned)cp, sr, sc);
        const int p = s_inv_r[sr], q = s_inv_c[sc];
        if (p < sR.nvalid() && q < sC.nvalid()) {
            const int si = sR.sector_of(p);
            const int sj = sM.mcol(si);
            if (sj >= 0 && q >= sC.sstart(sj) && q < sC.sstart(sj + 1))
                acc += dat[e] * __ldg(src + sM.moff(si) + (long long)(p - sR.sstart(si)) * sC.count(sj) + (q - sC.sstart(sj)));
        }
    }
    acc = block_sum(acc, red);
    if (tid == 0) {
        int t = 0;
        if (t1) t += s1 * t1[(long long)b * t1st];
        if (t2) t += s2 * t2[(long long)b * t2st];
        if (tsum) tsum[b] = t;
        // the result has no indexed edge: its single element exists iff the summed target vanishes
        int* Mrow = match + (long long)b * RT_MSTRIDE;
        Mrow[0] = t == 0 ? 2 : 0; Mrow[1] = 0; Mrow[2] = 0; Mrow[3] = t == 0 ? 2 : 0; Mrow[3 + RT_SMAX] = t == 0 ? 0 : -1;
        double* o = out + (long long)b * out_stride;
        o[0] = t == 0 ? acc : 0.0;
        o[1] = 0.0;
        if (stats) { atomicAdd(&stats[0], 2ull * (unsigned long long)total); atomicAdd(&stats[1], 2ull * (unsigned long long)total);
                     atomicAdd(&stats[2], 16ull * (unsigned long long)total); atomicAdd(&stats[8], 1ull); }
    }
}

// the binary search above lands on the LAST sector whose offset is <= e; sectors without partner share the offset of their
// successor, so the last one with that offset is the one that owns the element (or the final sentinel)

// ------------------------------------------------------------------------------------------------
// rt_gemm: C_s = A_s * B_s for every (chain, sector); 64 x 64 tiles, K slabs of 16, DMMA m8n8k4
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void rt_dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

constexpr int GT = 64;          // tile rows / columns
constexpr int GK = 16;          // K slab
constexpr int GLDA = GK + 4;    // conflict-free fragment reads
constexpr int GLDB = GT + 8;

__global__ void __launch_bounds__(128) rt_gemm_kernel(RtForm A, RtForm B, RtForm C, RtSpec spec, double* __restrict__ cdata, long long cstride,
                                                      int ksign, unsigned long long* stats) {
    __shared__ double As[GT * GLDA];
    __shared__ double Bs[GK * GLDB];
    __shared__ int hA[2][RT_HDR];       // A rows, A cols (k)
    __shared__ int hB[2][RT_HDR];       // B rows (k), B cols
    __shared__ int mA[RT_MSTRIDE], mB[RT_MSTRIDE], mC[RT_MSTRIDE];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < RT_HDR; i += 128) {
        hA[0][i] = A.rt[b * A.rts + i]; hA[1][i] = A.ct[b * A.cts + i];
        hB[0][i] = B.rt[b * B.rts + i]; hB[1][i] = B.ct[b * B.cts + i];
    }
    for (int i = tid; i < RT_MSTRIDE; i += 128) { mA[i] = A.match[b * A.mts + i]; mB[i] = B.match[b * B.mts + i]; }
    if (!spec.on)
        for (int i = tid; i < RT_MSTRIDE; i += 128) mC[i] = C.match[b * C.mts + i];
    __syncthreads();
    if (spec.on) rt_match_cta(hA[0], hB[1], spec, b, mC, blockIdx.x == 0);
    const RtTab aR(hA[0]), aK(hA[1]), bK(hB[0]), bN(hB[1]);
    const RtMatch MA(mA), MB(mB), MC(mC);
    const double* a = A.data + (long long)b * A.dstride;
    const double* bb = B.data + (long long)b * B.dstride;
    double* c = cdata + (long long)b * cstride;
    const int nsec = max(aR.nsec(), 0);
    const int g = lane >> 2, q = lane & 3;
    // flat tile index over (sector, tile_m, tile_n)
    int tile = blockIdx.x;
    int first = 0;      // tiles before the current sector
    for (int i = 0; i < nsec; ++i) {
        const int jc = MC.mcol(i);
        if (jc < 0) continue;
        const int m = aR.count(i), n = bN.count(jc);
        if (m == 0 || n == 0) continue;
        const int tm = (m + GT - 1) / GT, tn = (n + GT - 1) / GT;
        const int here = tm * tn;
        // operands of this sector
        int k = 0;
        long long aoff = 0, boff = 0;
        const int jk = MA.mcol(i);
        if (jk >= 0) {
            const int ib = bK.find(ksign * aK.skey(jk));
            if (ib >= 0 && MB.mcol(ib) == jc) {
                k = min(aK.count(jk), bK.count(ib));
                aoff = MA.moff(i);
                boff = MB.moff(ib);
            }
        }
        const long long coff = MC.moff(i);
        if (stats && blockIdx.x == 0 && tid == 0) {
            unsigned long long issued = 0;
            for (int t = 0; t < here; ++t) {
                const int rr = min(GT, m - (t / tn) * GT), cc = min(GT, n - (t % tn) * GT);
                issued += (unsigned long long)((rr + 15) / 16) * 2 * ((cc + 7) / 8) * ((k + 3) / 4);
            }
            atomicAdd(&stats[0], 2ull * m * n * k);
            atomicAdd(&stats[1], issued * 512ull);
            atomicAdd(&stats[2], 8ull * ((unsigned long long)m * k + (unsigned long long)k * n + (unsigned long long)m * n));
            atomicAdd(&stats[8], 1ull);
        }
        for (; tile < first + here; tile += gridDim.x) {
            const int t = tile - first;
            const int r0 = (t / tn) * GT, c0 = (t % tn) * GT;
            const int rows = min(GT, m - r0), cols = min(GT, n - c0);
            const int nfr = (cols + 7) >> 3;
            double acc[2][8][2];
#pragma unroll
            for (int x = 0; x < 2; ++x)
#pragma unroll
                for (int y = 0; y < 8; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
            for (int k0 = 0; k0 < k; k0 += GK) {
                const int kw = min(GK, k - k0);
                __syncthreads();
                // A tile: rows x kw (zero padded to GT x GK)
                for (int e = tid; e < GT * GK; e += 128) {
                    const int r = e / GK, kk = e - r * GK;
                    double v = 0.0;
                    if (r < rows && kk < kw) v = __ldg(a + aoff + (long long)(r0 + r) * k + k0 + kk);
                    As[r * GLDA + kk] = v;
                }
                for (int e = tid; e < GK * GT; e += 128) {
                    const int kk = e / GT, cc = e - kk * GT;
                    double v = 0.0;
                    if (kk < kw && cc < cols) v = __ldg(bb + boff + (long long)(k0 + kk) * n + c0 + cc);
                    Bs[kk * GLDB + cc] = v;
                }
                __syncthreads();
                if (warp * 16 < rows) {
#pragma unroll
                    for (int ks = 0; ks < GK; ks += 4) {
                        if (ks >= kw) break;
                        const double a0 = As[(warp * 16 + g) * GLDA + ks + q];
                        const double a1 = As[(warp * 16 + 8 + g) * GLDA + ks + q];
#pragma unroll
                        for (int y = 0; y < 8; ++y) {
                            if (y < nfr) {
                                const double bv = Bs[(ks + q) * GLDB + y * 8 + g];
                                rt_dmma(acc[0][y][0], acc[0][y][1], a0, bv);
                                rt_dmma(acc[1][y][0], acc[1][y][1], a1, bv);
                            }
                        }
                    }
                }
            }
            // epilogue: lane holds C[g][2q], C[g][2q+1] of every 8x8 fragment
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                const int r = warp * 16 + x * 8 + g;
                if (r < rows) {
#pragma unroll
                    for (int y = 0; y < 8; ++y) {
                        const int cc = y * 8 + 2 * q;
                        if (cc < cols) c[coff + (long long)(r0 + r) * n + c0 + cc] = acc[x][y][0];
                        if (cc + 1 < cols) c[coff + (long long)(r0 + r) * n + c0 + cc + 1] = acc[x][y][1];
                    }
                }
            }
            if (t == 0 && tid == 0 && ((m * n) & 1)) c[coff + (long long)m * n] = 0.0;     // alignment pad of an odd-sized sector
        }
        first += here;
    }
}

// ------------------------------------------------------------------------------------------------
// rt_gemm for SMALL sectors (cfg2: k, n of a sector <= ~40): one warp owns a 16-row x 32-column piece of one sector and runs on
// its own -- no shared-memory tiles, no block barriers.  Every lane loads its DMMA fragment elements straight from global memory:
// a sector matrix is contiguous and row-major, so the 8 x 4 A fragment of a warp-wide load covers 8 rows x 32 B of a k-wide row
// block (whole sectors for small k) and the B sector (<= 10 KiB) stays in L1.  HBM-bound by construction: A and C are streamed
// once, nothing else is touched (the tiled kernel above spent its time on header copies, zero-padded shared-memory tiles and two
// barriers per tile: 3 CTAs / SM, 0.25 TFLOP/s on cfg2).
// ------------------------------------------------------------------------------------------------
constexpr int kGemmDefault = 18;   // TNSP_RT_GEMM default (see tnsp_rt_gemm_f64)
constexpr int WR = 16;          // rows per warp item
constexpr int WC = 32;          // columns per warp item (4 DMMA fragments)

// per-sector descriptor built once per CTA in shared memory: the warps then find their items without touching the tables again
struct RtGemmDesc { int m, n, k, tn, start; long long aoff, boff, coff; };

template <int kGemmWarps, int WC = 32, int MINB = 32 / kGemmWarps>
__global__ void __launch_bounds__(kGemmWarps * 32, MINB) rt_gemm_warp_kernel(RtForm A, RtForm B, RtForm C, RtSpec spec, double* __restrict__ cdata,
                                                                          long long cstride, int ksign, unsigned long long* stats) {
    __shared__ int mC[RT_MSTRIDE];
    __shared__ RtGemmDesc dsc[RT_SMAX];
    __shared__ int n_desc, n_items;
    __shared__ int hh[4][RT_HDR];          // A rows, A cols (k), B rows (k), B cols: headers only (sector keys / starts)
    __shared__ int mAB[2][RT_MSTRIDE];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        // every table is read ONCE, all loads independent (one global-memory latency for the whole prologue)
        const int* src[4] = {A.rt + b * A.rts, A.ct + b * A.cts, B.rt + b * B.rts, B.ct + b * B.cts};
#pragma unroll
        for (int t = 0; t < 4; ++t)
            for (int i = tid; i < RT_HDR; i += kGemmWarps * 32) hh[t][i] = src[t][i];
        for (int i = tid; i < RT_MSTRIDE; i += kGemmWarps * 32) { mAB[0][i] = A.match[b * A.mts + i]; mAB[1][i] = B.match[b * B.mts + i]; }
        if (!spec.on)
            for (int i = tid; i < RT_MSTRIDE; i += kGemmWarps * 32) mC[i] = C.match[b * C.mts + i];
        __syncthreads();
    }
    if (spec.on) rt_match_cta(hh[0], hh[3], spec, b, mC, blockIdx.x == 0);
    {
        const RtTab aR(hh[0]), aK(hh[1]), bK(hh[2]), bN(hh[3]);
        const RtMatch MA(mAB[0]), MB(mAB[1]), MC(mC);
        const int nsec = max(aR.nsec(), 0);
        // one thread per row sector: operands and tile counts (every thread's loads are independent: one latency, not nsec)
        int here = 0;
        RtGemmDesc d;
        d.m = d.n = d.k = d.tn = d.start = 0; d.aoff = d.boff = d.coff = 0;
        if (tid < nsec) {
            const int i = tid;
            const int jc = MC.mcol(i);
            if (jc >= 0) {
                d.m = aR.count(i); d.n = bN.count(jc);
                if (d.m > 0 && d.n > 0) {
                    d.tn = (d.n + WC - 1) / WC;
                    here = ((d.m + WR - 1) / WR) * d.tn;
                    d.coff = MC.moff(i);
                    const int jk = MA.mcol(i);
                    if (jk >= 0) {
                        const int ib = bK.find(ksign * aK.skey(jk));
                        if (ib >= 0 && MB.mcol(ib) == jc) {
                            d.k = min(aK.count(jk), bK.count(ib));
                            d.aoff = MA.moff(i);
                            d.boff = MB.moff(ib);
                        }
                    }
                }
            }
        }
        // compact the non-empty sectors (order preserved) with a warp-level scan over <= 64 entries (two warps)
        __shared__ int cnt[2], tot[2];
        const unsigned has = __ballot_sync(0xffffffffu, here > 0);
        int incl = here;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (warp < 2 && lane == 31) { cnt[warp] = __popc(has); tot[warp] = incl; }
        __syncthreads();
        if (warp < 2 && here > 0) {
            const int slot = (warp ? cnt[0] : 0) + __popc(has & ((1u << lane) - 1u));
            d.start = (warp ? tot[0] : 0) + incl - here;
            dsc[slot] = d;
        }
        if (tid == 0) { n_desc = cnt[0] + cnt[1]; n_items = tot[0] + tot[1]; }
        __syncthreads();
        if (stats && blockIdx.x == 0 && tid < n_desc) {
            const RtGemmDesc& e = dsc[tid];
            const int tm = (e.m + WR - 1) / WR;
            unsigned long long issued = 0;
            for (int t = 0; t < e.tn; ++t) issued += (unsigned long long)tm * 2 * ((min(WC, e.n - t * WC) + 7) / 8) * ((e.k + 3) / 4);
            atomicAdd(&stats[0], 2ull * e.m * e.n * e.k);
            atomicAdd(&stats[1], issued * 512ull);
            atomicAdd(&stats[2], 8ull * ((unsigned long long)e.m * e.k + (unsigned long long)e.k * e.n + (unsigned long long)e.m * e.n));
            atomicAdd(&stats[8], 1ull);
        }
    }
    const double* a = A.data + (long long)b * A.dstride;
    const double* bb = B.data + (long long)b * B.dstride;
    double* c = cdata + (long long)b * cstride;
    const int g = lane >> 2, q = lane & 3;
    const int nd = n_desc, total = n_items;
    int cur = 0;
    for (int item = blockIdx.x * kGemmWarps + warp; item < total; item += gridDim.x * kGemmWarps) {
        while (cur + 1 < nd && dsc[cur + 1].start <= item) ++cur;
        const RtGemmDesc& e = dsc[cur];
        const int m = e.m, n = e.n, k = e.k;
        const int t = item - e.start;
        const int r0 = (t / e.tn) * WR, c0 = (t % e.tn) * WC;
        const int cols = min(WC, n - c0);
        const int nfr = (cols + 7) >> 3;
        constexpr int NF = WC / 8;
        double acc[2][NF][2];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < NF; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        const bool row0 = r0 + g < m, row1 = r0 + 8 + g < m;
        const double* ap0 = a + e.aoff + (long long)(r0 + g) * k + q;
        const double* ap1 = ap0 + 8ll * k;
        const double* bp = bb + e.boff + (long long)q * n + c0 + g;
        for (int ks = 0; ks < k; ks += 4) {
            const bool kin = ks + q < k;
            const double a0 = (kin && row0) ? __ldg(ap0 + ks) : 0.0;
            const double a1 = (kin && row1) ? __ldg(ap1 + ks) : 0.0;
#pragma unroll
            for (int y = 0; y < NF; ++y) {
                if (y < nfr) {
                    const double bv = (kin && y * 8 + g < cols) ? __ldg(bp + (long long)ks * n + y * 8) : 0.0;
                    rt_dmma(acc[0][y][0], acc[0][y][1], a0, bv);
                    rt_dmma(acc[1][y][0], acc[1][y][1], a1, bv);
                }
            }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int r = r0 + x * 8 + g;
            if (r < m) {
                double* cp = c + e.coff + (long long)r * n + c0;
#pragma unroll
                for (int y = 0; y < NF; ++y) {
                    const int cc = y * 8 + 2 * q;
                    if (cc < cols) cp[cc] = acc[x][y][0];
                    if (cc + 1 < cols) cp[cc + 1] = acc[x][y][1];
                }
            }
        }
        if (t == 0 && lane == 0 && ((m * n) & 1)) c[e.coff + (long long)m * n] = 0.0;     // alignment pad of an odd-sized sector
    }
}

// ------------------------------------------------------------------------------------------------
// rt_gemm, second generation of the small-sector kernel (the default): same work decomposition (one warp per 16 x 32 piece of one
// sector), two changes aimed at what ncu showed for the first one (a latency chain per item: one exposed global-memory round trip
// per k-step, 32 warps per SM, 10 % issue utilisation):
//   * B operand by TMA: the chain's whole B storage (all its sectors, back to back, even offsets) is brought into shared memory by
//     ONE bulk copy per CTA (cp.async.bulk + mbarrier, issued by one thread before the sector pairing is computed, so that it
//     overlaps the rest of the prologue); every B fragment then comes from shared memory instead of 4 L1 requests per k-step.
//     Falls back to global loads when the storage exceeds the buffer (cfg2: only the 216 x 216 operands) or is not 16-byte aligned;
//   * A operand software-pipelined ACROSS items: while a warp multiplies its current piece, the first 16 k of its NEXT piece are
//     already in flight (two pieces per warp in the memory system instead of one serialised k-loop; most cfg2 sectors have k <= 16).
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ unsigned rt_smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

struct RtGemmItem { int k, n, cols, rows0, rows1; long long aoff, coff; int boff, r0, c0, m, first; };

template <int kGemm2Warps, bool PREFETCH, int kGemmBsmem, int MINB>
__global__ void __launch_bounds__(kGemm2Warps * 32, MINB) rt_gemm_warp2_kernel(RtForm A, RtForm B, RtForm C, RtSpec spec, double* __restrict__ cdata,
                                                                            long long cstride, int ksign, unsigned long long* stats) {
    __shared__ __align__(16) double Bs[kGemmBsmem];
    __shared__ __align__(8) unsigned long long bbar;
    __shared__ int mC[RT_MSTRIDE];
    __shared__ RtGemmDesc dsc[RT_SMAX];
    __shared__ int n_desc, n_items;
    __shared__ int hh[4][RT_HDR];
    __shared__ int mAB[2][RT_MSTRIDE];
    constexpr int NT = kGemm2Warps * 32;
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* bb = B.data + (long long)b * B.dstride;
    {
        const int* src[4] = {A.rt + b * A.rts, A.ct + b * A.cts, B.rt + b * B.rts, B.ct + b * B.cts};
#pragma unroll
        for (int t = 0; t < 4; ++t)
            for (int i = tid; i < RT_HDR; i += NT) hh[t][i] = src[t][i];
        for (int i = tid; i < RT_MSTRIDE; i += NT) { mAB[0][i] = A.match[b * A.mts + i]; mAB[1][i] = B.match[b * B.mts + i]; }
        if (!spec.on)
            for (int i = tid; i < RT_MSTRIDE; i += NT) mC[i] = C.match[b * C.mts + i];
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(rt_smem_addr(&bbar)));
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
    }
    // B storage of this chain -> shared memory, one bulk copy (size = stored elements: every sector padded to an even count)
    const int bsize = mAB[1][0];
    const bool b_smem = bsize > 0 && bsize <= kGemmBsmem && ((reinterpret_cast<unsigned long long>(bb) & 15ull) == 0);
    if (b_smem && tid == 0) {
        const unsigned bytes = (unsigned)bsize * 8u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(rt_smem_addr(&bbar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(rt_smem_addr(Bs)), "l"(bb),
                     "r"(bytes), "r"(rt_smem_addr(&bbar))
                     : "memory");
    }
    if (spec.on) rt_match_cta(hh[0], hh[3], spec, b, mC, blockIdx.x == 0);
    {
        const RtTab aR(hh[0]), aK(hh[1]), bK(hh[2]), bN(hh[3]);
        const RtMatch MA(mAB[0]), MB(mAB[1]), MC(mC);
        const int nsec = max(aR.nsec(), 0);
        int here = 0;
        RtGemmDesc d;
        d.m = d.n = d.k = d.tn = d.start = 0; d.aoff = d.boff = d.coff = 0;
        if (tid < nsec) {
            const int i = tid;
            const int jc = MC.mcol(i);
            if (jc >= 0) {
                d.m = aR.count(i); d.n = bN.count(jc);
                if (d.m > 0 && d.n > 0) {
                    d.tn = (d.n + WC - 1) / WC;
                    here = ((d.m + WR - 1) / WR) * d.tn;
                    d.coff = MC.moff(i);
                    const int jk = MA.mcol(i);
                    if (jk >= 0) {
                        const int ib = bK.find(ksign * aK.skey(jk));
                        if (ib >= 0 && MB.mcol(ib) == jc) {
                            d.k = min(aK.count(jk), bK.count(ib));
                            d.aoff = MA.moff(i);
                            d.boff = MB.moff(ib);
                        }
                    }
                }
            }
        }
        __shared__ int cnt[2], tot[2];
        const unsigned has = __ballot_sync(0xffffffffu, here > 0);
        int incl = here;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
        if (warp < 2 && lane == 31) { cnt[warp] = __popc(has); tot[warp] = incl; }
        __syncthreads();
        if (warp < 2 && here > 0) {
            const int slot = (warp ? cnt[0] : 0) + __popc(has & ((1u << lane) - 1u));
            d.start = (warp ? tot[0] : 0) + incl - here;
            dsc[slot] = d;
        }
        if (tid == 0) { n_desc = cnt[0] + cnt[1]; n_items = tot[0] + tot[1]; }
        __syncthreads();
        if (stats && blockIdx.x == 0 && tid < n_desc) {
            const RtGemmDesc& e = dsc[tid];
            const int tm = (e.m + WR - 1) / WR;
            unsigned long long issued = 0;
            for (int t = 0; t < e.tn; ++t) issued += (unsigned long long)tm * 2 * ((min(WC, e.n - t * WC) + 7) / 8) * ((e.k + 3) / 4);
            atomicAdd(&stats[0], 2ull * e.m * e.n * e.k);
            atomicAdd(&stats[1], issued * 512ull);
            atomicAdd(&stats[2], 8ull * ((unsigned long long)e.m * e.k + (unsigned long long)e.k * e.n + (unsigned long long)e.m * e.n));
            atomicAdd(&stats[8], 1ull);
        }
    }
    const double* a = A.data + (long long)b * A.dstride;
    double* c = cdata + (long long)b * cstride;
    const int g = lane >> 2, q = lane & 3;
    const int nd = n_desc, total = n_items;
    const int stride = gridDim.x * kGemm2Warps;
    if (b_smem) {
        // all threads wait for the bulk copy (phase 0 of the barrier)
        unsigned done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(rt_smem_addr(&bbar)) : "memory");
    }
    // item -> descriptor (cursor only moves forward) and the first 16 k of its A rows
    auto locate = [&](int item, int& cur, RtGemmItem& it) {
        while (cur + 1 < nd && dsc[cur + 1].start <= item) ++cur;
        const RtGemmDesc& e = dsc[cur];
        const int t = item - e.start;
        it.m = e.m; it.n = e.n; it.k = e.k;
        it.r0 = (t / e.tn) * WR; it.c0 = (t % e.tn) * WC;
        it.cols = min(WC, e.n - it.c0);
        it.aoff = e.aoff; it.coff = e.coff; it.boff = (int)e.boff;
        it.rows0 = it.r0 + g < e.m; it.rows1 = it.r0 + 8 + g < e.m;
        it.first = t == 0;
    };
    auto load_a = [&](const RtGemmItem& it, int kc, double (&x0)[4], double (&x1)[4]) {
        const double* ap0 = a + it.aoff + (long long)(it.r0 + g) * it.k + q + kc;
        const double* ap1 = ap0 + 8ll * it.k;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool kin = kc + 4 * j + q < it.k;
            x0[j] = (kin && it.rows0) ? __ldg(ap0 + 4 * j) : 0.0;
            x1[j] = (kin && it.rows1) ? __ldg(ap1 + 4 * j) : 0.0;
        }
    };
    int item = blockIdx.x * kGemm2Warps + warp;
    int cur = 0, ncur = 0;
    RtGemmItem it;
    double a0[4], a1[4], n0[4], n1[4];
    bool have = item < total;
    if (have) { locate(item, cur, it); load_a(it, 0, a0, a1); ncur = cur; }
    while (have) {
        const int nitem = item + stride;
        const bool nhave = nitem < total;
        if (PREFETCH && nhave) { RtGemmItem nx; locate(nitem, ncur, nx); load_a(nx, 0, n0, n1); }       // in flight while the current piece is multiplied
        const int n = it.n, k = it.k, cols = it.cols;
        const int nfr = (cols + 7) >> 3;
        double acc[2][4][2];
#pragma unroll
        for (int x = 0; x < 2; ++x)
#pragma unroll
            for (int y = 0; y < 4; ++y) acc[x][y][0] = acc[x][y][1] = 0.0;
        for (int kc = 0; kc < k; kc += 16) {
            if (kc > 0) load_a(it, kc, a0, a1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int ks = kc + 4 * j;
                if (ks < k) {
                    const bool kin = ks + q < k;
                    double bv[4];
                    if (b_smem) {
                        const double* bp = Bs + it.boff + (ks + q) * n + it.c0 + g;
#pragma unroll
                        for (int y = 0; y < 4; ++y) bv[y] = (kin && y < nfr && y * 8 + g < cols) ? bp[y * 8] : 0.0;
                    } else {
                        const double* bp = bb + it.boff + (long long)(ks + q) * n + it.c0 + g;
#pragma unroll
                        for (int y = 0; y < 4; ++y) bv[y] = (kin && y < nfr && y * 8 + g < cols) ? __ldg(bp + y * 8) : 0.0;
                    }
#pragma unroll
                    for (int y = 0; y < 4; ++y) {
                        if (y < nfr) {
                            rt_dmma(acc[0][y][0], acc[0][y][1], a0[j], bv[y]);
                            rt_dmma(acc[1][y][0], acc[1][y][1], a1[j], bv[y]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int r = it.r0 + x * 8 + g;
            if (r < it.m) {
                double* cp = c + it.coff + (long long)r * n + it.c0;
#pragma unroll
                for (int y = 0; y < 4; ++y) {
                    const int cc = y * 8 + 2 * q;
                    if (cc < cols) cp[cc] = acc[x][y][0];
                    if (cc + 1 < cols) cp[cc + 1] = acc[x][y][1];
                }
            }
        }
        if (it.first && lane == 0 && ((it.m * n) & 1)) c[it.coff + (long long)it.m * n] = 0.0;     // alignment pad of an odd-sized sector
        item = nitem; have = nhave;
        if (have) {
            locate(item, cur, it);          // (descriptor re-read from shared memory: cheaper than keeping it in registers)
            if (PREFETCH) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { a0[j] = n0[j]; a1[j] = n1[j]; }
            } else load_a(it, 0, a0, a1);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// elementwise over the stored sectors (per-chain sizes from the match table)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rt_scale_kernel(const double* __restrict__ x, long long xs, const int* __restrict__ match, long long mts,
                                                       const double* __restrict__ alpha, int as, int op, double* __restrict__ y, long long ys) {
    const int b = blockIdx.y;
    const long long size = match[b * mts];
    const double a = alpha[(long long)b * as];
    const double f = op == 0 ? a : 1.0 / a;
    const double* X = x + b * xs;
    double* Y = y + b * ys;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < size; e += (long long)gridDim.x * 256)
        Y[e] = op == 0 ? X[e] * f : X[e] / a;
}

__global__ void __launch_bounds__(256) rt_binary_kernel(const double* __restrict__ x, long long xs, const double* __restrict__ w, long long ws,
                                                        const int* __restrict__ match, long long mts, int op, double* __restrict__ y, long long ys) {
    const int b = blockIdx.y;
    const long long size = match[b * mts];
    const double* X = x + b * xs;
    const double* W = w + b * ws;
    double* Y = y + b * ys;
    for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < size; e += (long long)gridDim.x * 256) {
        const double p = X[e], r = W[e];
        Y[e] = op == 0 ? p + r : op == 1 ? p - r : op == 2 ? p * r : (r == 0.0 ? p : p / r);
    }
}

__global__ void __launch_bounds__(256) rt_norm_kernel(const double* __restrict__ x, long long xs, const int* __restrict__ match, long long mts, int kind,
                                                      double* __restrict__ out) {
    __shared__ double red[32];
    const int b = blockIdx.x;
    const long long size = match[b * mts];
    const double* X = x + b * xs;
    double v = 0.0;
    for (long long e = threadIdx.x; e < size; e += 256) {
        const double t = X[e];
        if (kind == -1) v = fmax(v, fabs(t));
        else if (kind == 1) v += fabs(t);
        else v += t * t;
    }
    v = kind == -1 ? block_max(v, red) : block_sum(v, red);
    if (threadIdx.x == 0) out[b] = kind == 2 ? sqrt(v) : v;
}

__global__ void rt_scalar_kernel(const double* __restrict__ x, long long xs, const int* __restrict__ match, long long mts, double* __restrict__ out,
                                 int nb) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    out[b] = match[b * mts] > 0 ? x[b * xs] : 0.0;
}

}  // namespace tnsp

using namespace tnsp;

extern "C" int tnsp_rt_stats(int enable, uint64_t* out16, int reset) {
    // enable < 0: leave the switch alone; out16 != NULL: copy the counters to the host (synchronises); reset: zero them
    stats_alloc();
    if (!g_stats_dev) { set_error("tnsp_rt_stats: cudaMalloc"); return 1; }
    if (enable >= 0) g_stats_on = enable;
    if (out16 && cudaMemcpy(out16, g_stats_dev, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("tnsp_rt_stats: copy"); return 1; }
    if (reset == 1) cudaMemset(g_stats_dev, 0, 15 * sizeof(unsigned long long));      // slot 15 (overflows) is only cleared by reset == 2
    if (reset == 2) cudaMemset(g_stats_dev + 15, 0, sizeof(unsigned long long));
    return 0;
}

extern "C" int tnsp_rt_sort_i32(int n_edges, const int32_t* const* labels, const int64_t* lstrides, const int32_t* dims, const int32_t* signs,
                                int64_t M, int32_t* table, int nbT, void* stream) {
    if (n_edges > 8) { set_error("tnsp_rt_sort_i32: at most 8 edges per group"); return 1; }
    if (M >= (1ll << 31)) { set_error("tnsp_rt_sort_i32: merged dimension beyond int32"); return 1; }
    RtEdges E;
    E.n = n_edges;
    for (int i = 0; i < n_edges; ++i) { E.lab[i] = labels[i]; E.lstride[i] = lstrides[i]; E.dim[i] = dims[i]; E.sign[i] = signs[i]; }
    if (nbT == 0) return 0;
    if (M <= kSortWarpMax)
        rt_sort_warp_kernel<<<(nbT + kSortWarpPerCta - 1) / kSortWarpPerCta, 32 * kSortWarpPerCta, 0, (cudaStream_t)stream>>>(E, (int)M, table,
                                                                                                                    RT_HDR + 2 * M, nbT);
    else
        rt_sort_kernel<256><<<nbT, 256, 0, (cudaStream_t)stream>>>(E, (int)M, table, RT_HDR + 2 * M);
    return check_launch("tnsp_rt_sort_i32");
}

extern "C" int tnsp_rt_match_i32(const int32_t* rt, int64_t rt_stride, int rs, const int32_t* ct, int64_t ct_stride, int cs, const int32_t* t1,
                                 int t1_stride, int s1, const int32_t* t2, int t2_stride, int s2, int32_t* match, int32_t* tsum, int nbm,
                                 int64_t cap, void* stream) {
    if (nbm == 0) return 0;
    rt_match_kernel<<<(nbm + 3) / 4, 128, 0, (cudaStream_t)stream>>>(rt, rt_stride, rs, ct, ct_stride, cs, t1, t1_stride, s1, t2, t2_stride, s2, match,
                                                                     tsum, nbm, cap, rt_overflow_ptr());
    return check_launch("tnsp_rt_match_i32");
}

// CTA size of the tile regrouping kernels (TNSP_RT_REPACK_THREADS = 256 | 128 | 64 overrides the default)
// Measured on cfg2 (B200, 2368 chains, ms per 3 steps): single regrouping 401 / 305 / 368 at 256 / 128 / 64 threads, the pair kernel
// 510 / 455 / 422 -- a chain regroups ~1.7 k elements, so CTAs in flight per SM matter more than threads per CTA.
static int repack_threads(bool pair = false) {
    static int v[2] = {0, 0};
    if (!v[0]) {
        const char* e = getenv("TNSP_RT_REPACK_THREADS");
        const int x = e ? atoi(e) : 0;
        const bool ok = x == 256 || x == 128 || x == 64;
        v[0] = ok ? x : kRepackTileThreads;
        v[1] = ok ? x : kRepackPairThreads;
    }
    return v[pair ? 1 : 0];
}

extern "C" int tnsp_rt_match_multi_i32(const tnsp_rt_match_job* jobs, int n_jobs, int nbm, void* stream) {
    if (nbm == 0 || n_jobs == 0) return 0;
    if (n_jobs > 4) { set_error("tnsp_rt_match_multi_i32: at most 4 pairings per launch"); return 1; }
    RtMatchJobs J;
    for (int i = 0; i < 4; ++i) J.j[i] = jobs[i < n_jobs ? i : 0];
    rt_match_multi_kernel<<<dim3((nbm + 3) / 4, n_jobs), 128, 0, (cudaStream_t)stream>>>(J, nbm, rt_overflow_ptr());
    return check_launch("tnsp_rt_match_multi_i32");
}

// tiles of the densest possible layout / 2 (a CTA loops over its tiles; sectors only cover a fraction of the M x N index space)
static dim3 tile_grid(int64_t M, int64_t N, int nb, int nz) {
    int64_t gx = (M * N / 2048 + 7) / 8;
    if (gx < 1) gx = 1;
    if (gx > 512) gx = 512;
    return dim3((unsigned)gx, (unsigned)nb, (unsigned)nz);
}

static dim3 repack_grid(int64_t work, int nb, int nz) {
    int64_t gx = (work + kRepackThreads * 4 - 1) / (kRepackThreads * 4);
    if (gx < 1) gx = 1;
    if (gx > 4096) gx = 4096;
    return dim3((unsigned)gx, (unsigned)nb, (unsigned)nz);
}

extern "C" int tnsp_rt_repack_f64(const int32_t* plan, const tnsp_rt_form* src, const tnsp_rt_form* dst, const tnsp_rt_match_spec* dst_match,
                                  double* dst_data, int64_t dst_stride, int64_t work, int nb, void* stream) {
    // src->rt == NULL: the source is a dense array (src->data, src->data_stride); dst->rt == NULL: the destination is dense with
    // `work` elements per chain; otherwise `work` is an upper bound of the stored elements per chain (sizes the grid only)
    if (nb == 0) return 0;
    const bool sd = src->rt == nullptr, dd = dst->rt == nullptr;
    if (sd && dd) { set_error("tnsp_rt_repack_f64: one side must be sector-compact"); return 1; }
    RtRepackArgs p;
    p.plan = plan; p.S = to_form(src); p.D = to_form(dst); p.spec = to_spec(dst_match); p.dst = dst_data; p.dst_stride = dst_stride;
    p.dense_size = dd ? work : 0;
    const dim3 grid = repack_grid(work, nb, 1);
    cudaStream_t st = (cudaStream_t)stream;
    if (sd) rt_repack_kernel<true, false><<<grid, kRepackThreads, 0, st>>>(p, rt_stats_ptr());
    else if (dd) rt_repack_kernel<false, true><<<grid, kRepackThreads, 0, st>>>(p, rt_stats_ptr());
    else if (repack_threads() == 128) rt_repack_tile_kernel<128><<<tile_grid(dst->M, dst->N, nb, 1), 128, 0, st>>>(p, rt_stats_ptr());
    else if (repack_threads() == 64) rt_repack_tile_kernel<64><<<tile_grid(dst->M, dst->N, nb, 1), 64, 0, st>>>(p, rt_stats_ptr());
    else rt_repack_tile_kernel<256><<<tile_grid(dst->M, dst->N, nb, 1), 256, 0, st>>>(p, rt_stats_ptr());
    return check_launch("tnsp_rt_repack_f64");
}

extern "C" int tnsp_rt_repack_signed_f64(const int32_t* plan, const tnsp_rt_form* src, const tnsp_rt_form* dst, const tnsp_rt_match_spec* dst_match,
                                         double* dst_data, int64_t dst_stride, const int32_t* quad, const int32_t* per_chain,
                                         int64_t per_chain_stride, int n_entries, const int32_t* const* labels, const int64_t* lstrides,
                                         int fermi_mask, int nb, void* stream) {
    if (nb == 0) return 0;
    if (!src->rt || !dst->rt) { set_error("tnsp_rt_repack_signed_f64: both layouts must be sector-compact"); return 1; }
    if (n_entries > 24) { set_error("tnsp_rt_repack_signed_f64: at most 24 indexed edges"); return 1; }
    RtRepackArgs p;
    p.plan = plan; p.S = to_form(src); p.D = to_form(dst); p.spec = to_spec(dst_match); p.dst = dst_data; p.dst_stride = dst_stride; p.dense_size = 0;
    RtSign sg;
    sg.quad = quad; sg.per_chain = per_chain; sg.pcs = per_chain_stride; sg.fermi = fermi_mask;
    for (int i = 0; i < 24; ++i) { sg.lab[i] = i < n_entries ? labels[i] : nullptr; sg.lst[i] = i < n_entries ? lstrides[i] : 0; }
    if (repack_threads() == 128) rt_repack_signed_kernel<128><<<tile_grid(dst->M, dst->N, nb, 1), 128, 0, (cudaStream_t)stream>>>(p, sg, rt_stats_ptr());
    else if (repack_threads() == 64) rt_repack_signed_kernel<64><<<tile_grid(dst->M, dst->N, nb, 1), 64, 0, (cudaStream_t)stream>>>(p, sg, rt_stats_ptr());
    else rt_repack_signed_kernel<256><<<tile_grid(dst->M, dst->N, nb, 1), 256, 0, (cudaStream_t)stream>>>(p, sg, rt_stats_ptr());
    return check_launch("tnsp_rt_repack_signed_f64");
}

extern "C" int tnsp_rt_repack_pair_f64(const int32_t* plan0, const tnsp_rt_form* src0, const tnsp_rt_form* dst0, const tnsp_rt_match_spec* match0,
                                       double* dst_data0, int64_t dst_stride0, int64_t work0, const int32_t* plan1, const tnsp_rt_form* src1,
                                       const tnsp_rt_form* dst1, const tnsp_rt_match_spec* match1, double* dst_data1, int64_t dst_stride1,
                                       int64_t work1, int nb, void* stream) {
    if (nb == 0) return 0;
    if (!src0->rt || !dst0->rt || !src1->rt || !dst1->rt) { set_error("tnsp_rt_repack_pair_f64: both regroupings must be sector-compact"); return 1; }
    RtRepackPair p;
    p.a[0].plan = plan0; p.a[0].S = to_form(src0); p.a[0].D = to_form(dst0); p.a[0].spec = to_spec(match0); p.a[0].dst = dst_data0;
    p.a[0].dst_stride = dst_stride0; p.a[0].dense_size = 0;
    p.a[1].plan = plan1; p.a[1].S = to_form(src1); p.a[1].D = to_form(dst1); p.a[1].spec = to_spec(match1); p.a[1].dst = dst_data1;
    p.a[1].dst_stride = dst_stride1; p.a[1].dense_size = 0;
    const dim3 g0 = tile_grid(dst0->M, dst0->N, nb, 2), g1 = tile_grid(dst1->M, dst1->N, nb, 2);
    const dim3 gp(g0.x > g1.x ? g0.x : g1.x, nb, 2);
    if (repack_threads(true) == 128) rt_repack_tile_pair_kernel<128><<<gp, 128, 0, (cudaStream_t)stream>>>(p, rt_stats_ptr());
    else if (repack_threads(true) == 64) rt_repack_tile_pair_kernel<64><<<gp, 64, 0, (cudaStream_t)stream>>>(p, rt_stats_ptr());
    else rt_repack_tile_pair_kernel<256><<<gp, 256, 0, (cudaStream_t)stream>>>(p, rt_stats_ptr());
    return check_launch("tnsp_rt_repack_pair_f64");
}

extern "C" int tnsp_rt_gemm_f64(const tnsp_rt_form* a, const tnsp_rt_form* b, const tnsp_rt_form* c, const tnsp_rt_match_spec* c_match,
                                double* c_data, int64_t c_stride, int ksign, int nb, void* stream) {
    if (nb == 0) return 0;
    if (a->N <= 512 && b->N <= 512) {
        // small sectors: warp-autonomous kernel; the grid covers the most items a chain can have (every sector adds at most one
        // partial piece per direction), spare CTAs leave at once
        int64_t items = ((a->M + WR - 1) / WR + RT_SMAX / 4) * ((b->N + 32 - 1) / 32);
        // TNSP_RT_GEMM = <generation><warps>: 18 / 14 / 12 first generation with 8 / 4 / 2 warps per CTA; 38 = 8 warps, 16-column pieces; 24 = TMA-fed B (16 KiB) without the
        // cross-item prefetch, 4 warps; 25 = TMA-fed B (32 KiB) + prefetch, 4 warps
        static int gen = -1;
        if (gen < 0) { const char* e = getenv("TNSP_RT_GEMM"); gen = e ? atoi(e) : kGemmDefault; }
        cudaStream_t st = (cudaStream_t)stream;
        const RtForm fa = to_form(a), fb = to_form(b), fc = to_form(c);
        const RtSpec sp = to_spec(c_match);
        auto grid = [&](int warps) { int64_t gx = (items + 8 * warps - 1) / (8 * warps); if (gx > 256) gx = 256; return dim3((unsigned)gx, (unsigned)nb); };
        switch (gen) {
        case 25: rt_gemm_warp2_kernel<4, true, 4096, 5><<<grid(4), 128, 0, st>>>(fa, fb, fc, sp, c_data, c_stride, ksign, rt_stats_ptr()); break;
        case 24: rt_gemm_warp2_kernel<4, false, 2048, 8><<<grid(4), 128, 0, st>>>(fa, fb, fc, sp, c_data, c_stride, ksign, rt_stats_ptr()); break;
        case 14: rt_gemm_warp_kernel<4><<<grid(4), 128, 0, st>>>(fa, fb, fc, sp, c_data, c_stride, ksign, rt_stats_ptr()); break;
        case 12: rt_gemm_warp_kernel<2><<<grid(2), 64, 0, st>>>(fa, fb, fc, sp, c_data, c_stride, ksign, rt_stats_ptr()); break;
        case 38: items = ((a->M + WR - 1) / WR + RT_SMAX / 4) * ((b->N + 15) / 16);     // 16-column pieces: 8 accumulators, 5 CTAs per SM
                 rt_gemm_warp_kernel<8, 16, 5><<<grid(8), 256, 0, st>>>(fa, fb, fc, sp, c_data, c_stride, ksign, rt_stats_ptr()); break;
        default: rt_gemm_warp_kernel<8><<<grid(8), 256, 0, st>>>(fa, fb, fc, sp, c_data, c_stride, ksign, rt_stats_ptr()); break;
        }
        return check_launch("tnsp_rt_gemm_f64(warp)");
    }
    int64_t tiles = ((a->M + GT - 1) / GT) * ((b->N + GT - 1) / GT);
    if (tiles < 1) tiles = 1;
    if (tiles > 64) tiles = 64;
    dim3 grid((unsigned)tiles, (unsigned)nb);
    rt_gemm_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(to_form(a), to_form(b), to_form(c), to_spec(c_match), c_data, c_stride, ksign,
                                                           rt_stats_ptr());
    return check_launch("tnsp_rt_gemm_f64");
}

extern "C" int tnsp_rt_dot_f64(const int32_t* plan, const tnsp_rt_form* src, const tnsp_rt_form* dst, const int32_t* t1, int t1_stride, int s1,
                               const int32_t* t2, int t2_stride, int s2, double* out, int64_t out_stride, int32_t* match, int32_t* tsum, int nb,
                               void* stream) {
    if (nb == 0) return 0;
    rt_dot_kernel<<<nb, kDotThreads, 0, (cudaStream_t)stream>>>(plan, to_form(src), to_form(dst), t1, t1_stride, s1, t2, t2_stride, s2, out, out_stride,
                                                                match, tsum, rt_stats_ptr());
    return check_launch("tnsp_rt_dot_f64");
}

extern "C" int tnsp_rt_scale_f64(const double* x, int64_t x_stride, const int32_t* match, int64_t match_stride, const double* alpha, int alpha_stride,
                                 int op, double* y, int64_t y_stride, int64_t cap, int nb, void* stream) {
    if (nb == 0) return 0;
    int64_t gx = (cap + 1023) / 1024;
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    rt_scale_kernel<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(x, x_stride, match, match_stride, alpha, alpha_stride, op, y, y_stride);
    return check_launch("tnsp_rt_scale_f64");
}

extern "C" int tnsp_rt_binary_f64(const double* x, int64_t x_stride, const double* w, int64_t w_stride, const int32_t* match, int64_t match_stride,
                                  int op, double* y, int64_t y_stride, int64_t cap, int nb, void* stream) {
    if (nb == 0) return 0;
    int64_t gx = (cap + 1023) / 1024;
    gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
    rt_binary_kernel<<<dim3((unsigned)gx, (unsigned)nb), 256, 0, (cudaStream_t)stream>>>(x, x_stride, w, w_stride, match, match_stride, op, y, y_stride);
    return check_launch("tnsp_rt_binary_f64");
}

extern "C" int tnsp_rt_norm_f64(const double* x, int64_t x_stride, const int32_t* match, int64_t match_stride, int kind, double* out, int nb,
                                void* stream) {
    if (nb == 0) return 0;
    rt_norm_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(x, x_stride, match, match_stride, kind, out);
    return check_launch("tnsp_rt_norm_f64");
}

extern "C" int tnsp_rt_scalar_f64(const double* x, int64_t x_stride, const int32_t* match, int64_t match_stride, double* out, int nb, void* stream) {
    if (nb == 0) return 0;
    rt_scalar_kernel<<<(nb + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, x_stride, match, match_stride, out, nb);
    return check_launch("tnsp_rt_scalar_f64");
}
