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

inline RtForm to_form(const tnsp_rt_form* f) {
    RtForm r;
    r.data = f->data; r.dstride = f->data_stride;
    r.rt = f->rt; r.rts = f->rt_stride; r.M = f->M;
    r.ct = f->ct; r.cts = f->ct_stride; r.N = f->N;
    r.match = f->match; r.mts = f->match_stride;
    return r;
}


}  // namespace tnsp
