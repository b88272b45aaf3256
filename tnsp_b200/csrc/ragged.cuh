// Shared device-side views of the sector-compact tables (tnsp_b200/TAT/ragged.py; layouts documented there).
#pragma once
#include "common.cuh"
#include <climits>

namespace tnsp {

constexpr int RT_SMAX = 64;
constexpr int RT_HDR = 3 + 2 * RT_SMAX;
constexpr int RT_MSTRIDE = 4 + 2 * RT_SMAX;
constexpr int RT_DEAD_MIN = 1 << 29;
constexpr int RT_EMPTY = INT_MIN;

// work counters of the instrumented bench pass (bench.py roofline): 0 gemm algorithmic flops (sum 2 m n k over the sectors),
// 1 gemm executed flops (DMMA.8x8x4 issued x 512), 2 gemm algorithmic bytes 8 (mk + kn + mn), 3 repack elements, 4 qr bytes
// 8 (2mn + mk + kn), 5 qr flops 4 m n k, 6 svd bytes 8 (mn + mk + kn + k), 7 sectors factorised, 8 gemm sectors
unsigned long long* rt_stats_ptr();
unsigned long long* rt_overflow_ptr();    // device counter of chains dropped because their sectors did not fit the destination buffer      // device pointer to the 16 counters, nullptr while the counters are switched off

// decoded view of one chain's group table
struct RtTab {
    const int* p;
    __device__ RtTab(const int* base) : p(base) {}
    __device__ int nsec() const { return p[0]; }
    __device__ int nvalid() const { return p[1]; }
    __device__ int skey(int i) const { return p[2 + i]; }
    __device__ int sstart(int i) const { return p[2 + RT_SMAX + i]; }
    __device__ int count(int i) const { return p[3 + RT_SMAX + i] - p[2 + RT_SMAX + i]; }
    __device__ int find(int key) const {
        const int n = p[0];
        for (int i = 0; i < n; ++i)
            if (p[2 + i] == key) return i;
        return -1;
    }
    // sector holding sorted position pos (pos < nvalid)
    __device__ int sector_of(int pos) const {
        int lo = 0, hi = p[0];   // sstart[lo] <= pos < sstart[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (p[2 + RT_SMAX + mid] <= pos) lo = mid; else hi = mid;
        }
        return lo;
    }
};
struct RtMatch {
    const int* p;
    __device__ RtMatch(const int* base) : p(base) {}
    __device__ int size() const { return p[0]; }
    __device__ int moff(int i) const { return p[2 + i]; }
    __device__ int mcol(int i) const { return p[3 + RT_SMAX + i]; }
};

struct RtForm {
    const double* data; long long dstride;
    const int* rt; long long rts; long long M;
    const int* ct; long long cts; long long N;
    const int* match; long long mts;
};

// sector pairing computed inside a consumer kernel (every CTA of the chain recomputes it in shared memory, CTA x == 0 stores it)
struct RtSpec {
    int on;                 // 0: the form's match table is read from memory
    int rs, cs;
    const int* t1; int t1st, s1;
    const int* t2; int t2st, s2;
    int* out; long long outs;
    int* tsum;
    long long cap;                  // elements the destination buffer holds per chain (0: not checked)
    unsigned long long* flag;       // overflow counter: a chain whose sectors do not fit is stored EMPTY and counted here
};
inline RtSpec to_spec(const tnsp_rt_match_spec* p) {
    RtSpec r;
    r.on = p != nullptr;
    if (p) { r.rs = p->rs; r.cs = p->cs; r.t1 = p->t1; r.t1st = p->t1_stride; r.s1 = p->s1; r.t2 = p->t2; r.t2st = p->t2_stride; r.s2 = p->s2;
             r.out = p->match_out; r.outs = p->match_out_stride; r.tsum = p->tsum_out; r.cap = p->cap; r.flag = rt_overflow_ptr(); }
    else { r.rs = r.cs = 1; r.t1 = r.t2 = nullptr; r.t1st = r.t2st = r.s1 = r.s2 = 0; r.out = nullptr; r.outs = 0; r.tsum = nullptr; r.cap = 0; r.flag = nullptr; }
    return r;
}
// all threads of the CTA; hR / hC: table headers in shared memory; m: MSTRIDE ints of shared memory
__device__ __forceinline__ void rt_match_cta(const int* hR, const int* hC, const RtSpec& sp, int b, int* m, bool store) {
    const int tid = threadIdx.x, nt = blockDim.x;
    int t = 0;
    if (sp.t1) t += sp.s1 * sp.t1[(long long)b * sp.t1st];
    if (sp.t2) t += sp.s2 * sp.t2[(long long)b * sp.t2st];
    const int nr = max(hR[0], 0), nc = max(hC[0], 0);
    for (int i = tid; i < nr; i += nt) {
        const int want = t - sp.rs * hR[2 + i];
        int j = -1;
        for (int jj = 0; jj < nc; ++jj)
            if (sp.cs * hC[2 + jj] == want) { j = jj; break; }
        int sz = 0;
        if (j >= 0) { sz = (hR[3 + RT_SMAX + i] - hR[2 + RT_SMAX + i]) * (hC[3 + RT_SMAX + j] - hC[2 + RT_SMAX + j]); sz += sz & 1; }
        m[3 + RT_SMAX + i] = j;
        m[2 + i] = sz;
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int i = 0; i < nr; ++i) { const int sz = m[2 + i]; m[2 + i] = acc; acc += sz; }
        m[2 + nr] = acc;
        bool bad = hR[0] < 0 || hC[0] < 0;
        if (sp.cap > 0 && acc > sp.cap) {            // learnt capacity exceeded: the chain is stored empty, loudly (backend.rt_overflow)
            bad = true;
            if (store && sp.flag) atomicAdd(sp.flag, 1ull);
            for (int i = 0; i <= nr; ++i) m[2 + i] = 0;
            for (int i = 0; i < nr; ++i) m[3 + RT_SMAX + i] = -1;
        }
        m[0] = bad ? 0 : acc;
        m[1] = bad ? 1 : 0;
        if (store && sp.tsum) sp.tsum[sp.outs ? b : 0] = t;     // one row for all chains when the pairing is chain-independent
    }
    __syncthreads();
    if (store && sp.out) {
        int* o = sp.out + (long long)b * sp.outs;
        for (int i = tid; i < 3 + nr; i += nt) o[i] = m[i];
        for (int i = tid; i < nr; i += nt) o[3 + RT_SMAX + i] = m[3 + RT_SMAX + i];
    }
}

inline RtForm to_form(const tnsp_rt_form* f) {
    RtForm r;
    r.data = f->data; r.dstride = f->data_stride;
    r.rt = f->rt; r.rts = f->rt_stride; r.M = f->M;
    r.ct = f->ct; r.cts = f->ct_stride; r.N = f->N;
    r.match = f->match; r.mts = f->match_stride;
    return r;
}


}  // namespace tnsp
