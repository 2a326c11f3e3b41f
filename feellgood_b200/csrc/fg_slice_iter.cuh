// Row ownership of the Krylov kernels (which SELL slices a warp works on), as host + device code so that
// tests/cpp/slice_iter_test.cu can enumerate it on the CPU.
#pragma once
#include <climits>

#if defined(__CUDACC__)
#define FG_SI_HD __host__ __device__ __forceinline__
#else
#define FG_SI_HD inline
#endif

namespace fg
{
// Which slices a warp owns.  Regular rounds: slice g + k W for k = 0 .. (W = warps of the grid, g = this
// warp's index in the grid), so that at any moment the whole grid works on one contiguous front of W slices
// (the gathered images of a front are shared through L2 and, inside a CTA, through L1).  The last,
// incomplete round is dealt out per CTA instead (`tail`: at most one extra slice per warp), so that every SM
// ends with the same number of slices to one, whatever the ratio of slices to warps (it is ~4 per warp on
// the 8-GPU partition of the 20 M-tet mesh, where a round-robin tail left 17 SMs with 25 % more work).
struct SliceIter
    {
    int first;     // g, or INT_MAX when the warp has no regular slice
    int W;         // stride of the regular rounds
    int main_end;  // regular slices are below this index
    int tail;      // the warp's slice of the last round, or INT_MAX
    FG_SI_HD int begin() const { return first < main_end ? first : tail; }
    FG_SI_HD int next(int s) const
        {
        if (s >= main_end) return INT_MAX;  // s was the tail slice
        const int n = s + W;
        return n < main_end ? n : tail;
        }
    };
// plain grid-stride ownership (stand-alone kernels)
FG_SI_HD SliceIter slices_strided(int g, int W, int nslice)
    {
    SliceIter it;
    it.first = g;
    it.W = W;
    it.main_end = nslice;
    it.tail = INT_MAX;
    return it;
    }
// front + per-CTA tail (persistent kernel): n units (slices or gather blocks) dealt to `per_cta` owners per
// CTA (warps or thread groups), this owner being number `id` of CTA `cta` of `nctas`
FG_SI_HD SliceIter slices_balanced_at(int n, int per_cta, int id, int cta, int nctas)
    {
    const int W = nctas * per_cta;
    const int q = n / W, rem = n - q * W;
    SliceIter it;
    it.first = cta * per_cta + id;
    it.W = W;
    it.main_end = q * W;
    const int t0 = it.main_end + (int)(((long long)rem * cta) / nctas);
    const int t1 = it.main_end + (int)(((long long)rem * (cta + 1)) / nctas);
    it.tail = t0 + id < t1 ? t0 + id : INT_MAX;
    return it;
    }
}  // namespace fg
