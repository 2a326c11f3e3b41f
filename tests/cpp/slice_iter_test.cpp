// Row ownership of the persistent solve kernel, enumerated on the CPU (g++, no GPU): the very
// SliceIter / slices_balanced_at / slices_strided the kernels call (feellgood_b200/csrc/fg_slice_iter.cuh).
//   * every slice is owned by exactly one warp, and a warp meets its slices in increasing order;
//   * regular rounds form a front: round k of warp g is slice g + k W;
//   * the incomplete last round is dealt out per CTA: contiguous, at most one slice per warp, and the CTA
//     totals differ by at most one slice whatever the ratio of slices to warps.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fg_slice_iter.cuh"

static int fails = 0;
#define CHECK(c, ...)                                                                                  \
    do                                                                                                 \
        {                                                                                              \
        if (!(c))                                                                                      \
            {                                                                                          \
            if (fails++ < 20)                                                                          \
                {                                                                                      \
                std::fprintf(stderr, "FAIL %s:%d %s: ", __FILE__, __LINE__, #c);                       \
                std::fprintf(stderr, __VA_ARGS__);                                                     \
                std::fprintf(stderr, "\n");                                                            \
                }                                                                                      \
            }                                                                                          \
        }                                                                                              \
    while (0)

static void balanced(int n, int nctas, int per_cta)
    {
    std::vector<int> owner(n, -1);
    const int W = nctas * per_cta, q = n / W, rem = n - q * W;
    int cta_min = 1 << 30, cta_max = -1, prev_tail_end = q * W;
    for (int c = 0; c < nctas; c++)
        {
        int cta_count = 0, tail_lo = 1 << 30, tail_hi = -1, tails = 0;
        for (int id = 0; id < per_cta; id++)
            {
            const fg::SliceIter it = fg::slices_balanced_at(n, per_cta, id, c, nctas);
            int k = 0, last = -1, count = 0;
            for (int s = it.begin(); s < n; s = it.next(s))
                {
                CHECK(s > last, "order n=%d ctas=%d per=%d cta=%d id=%d s=%d last=%d", n, nctas, per_cta, c, id, s, last);
                CHECK(s >= 0 && owner[s] == -1, "owned twice n=%d ctas=%d per=%d s=%d", n, nctas, per_cta, s);
                if (s >= 0 && s < n) owner[s] = c * per_cta + id;
                if (s < q * W)
                    CHECK(s == c * per_cta + id + k * W, "front n=%d ctas=%d per=%d s=%d k=%d", n, nctas, per_cta, s, k);
                else
                    {
                    tails++;
                    if (s < tail_lo) tail_lo = s;
                    if (s > tail_hi) tail_hi = s;
                    }
                last = s;
                k++;
                count++;
                if (count > q + 1) break;
                }
            CHECK(count == q || count == q + 1, "count n=%d ctas=%d per=%d cta=%d id=%d count=%d q=%d", n, nctas, per_cta, c, id, count, q);
            cta_count += count;
            }
        if (tails > 0)
            {  // the CTA's share of the last round is one contiguous range that starts where the previous ended
            CHECK(tail_hi - tail_lo + 1 == tails, "tail not contiguous n=%d ctas=%d per=%d cta=%d", n, nctas, per_cta, c);
            CHECK(tail_lo == prev_tail_end, "tail gap n=%d ctas=%d per=%d cta=%d", n, nctas, per_cta, c);
            prev_tail_end = tail_hi + 1;
            }
        CHECK(tails <= per_cta, "tail larger than the CTA n=%d ctas=%d per=%d", n, nctas, per_cta);
        if (cta_count < cta_min) cta_min = cta_count;
        if (cta_count > cta_max) cta_max = cta_count;
        }
    CHECK(prev_tail_end == n, "tail end n=%d ctas=%d per=%d end=%d", n, nctas, per_cta, prev_tail_end);
    CHECK(cta_max - cta_min <= 1, "imbalance n=%d ctas=%d per=%d min=%d max=%d rem=%d", n, nctas, per_cta, cta_min, cta_max, rem);
    for (int s = 0; s < n; s++) CHECK(owner[s] >= 0, "unowned n=%d ctas=%d per=%d s=%d", n, nctas, per_cta, s);
    }

static void strided(int n, int W)
    {
    std::vector<int> owner(n, -1);
    for (int g = 0; g < W; g++)
        {
        const fg::SliceIter it = fg::slices_strided(g, W, n);
        int k = 0;
        for (int s = it.begin(); s < n; s = it.next(s), k++)
            {
            CHECK(s == g + k * W && owner[s] == -1, "strided n=%d W=%d g=%d s=%d", n, W, g, s);
            owner[s] = g;
            }
        }
    for (int s = 0; s < n; s++) CHECK(owner[s] >= 0, "strided unowned n=%d W=%d s=%d", n, W, s);
    }

int main()
    {
    // the shapes of the five configurations and of the 2/4/8-GPU partitions of the 20 M-tet mesh
    // (fg_solver_launch_shape), then a sweep over ratios of slices to warps around every boundary
    const int shapes[][3] = {{6, 1, 8},      {1483, 93, 16},  {5968, 148, 24},  {29541, 148, 29}, {156252, 148, 32},
                             {78126, 148, 32}, {39035, 148, 30}, {39091, 148, 30}, {19517, 148, 27}, {19574, 148, 27},
                             {3034, 380, 8},  {1827, 115, 16}, {97, 7, 16}};
    for (const auto &sh : shapes) balanced(sh[0], sh[1], sh[2]);
    const int grids[] = {1, 2, 7, 93, 148, 592}, pers[] = {1, 2, 8, 16, 27, 32};
    for (int nctas : grids)
        for (int per : pers)
            {
            const int W = nctas * per;
            const int ns[] = {1,     2,         W - 1,     W,         W + 1,         2 * W - 1, 2 * W + nctas - 1,
                              3 * W, 3 * W + 1, 4 * W + W / 2, 5 * W - 1, 17 * W + 5, nctas,     nctas + 1};
            for (int n : ns)
                if (n >= 1) balanced(n, nctas, per);
            }
    unsigned int x = 12345u;
    for (int t = 0; t < 400; t++)
        {  // random shapes
        x = x * 1664525u + 1013904223u;
        const int nctas = 1 + (int)((x >> 8) % 200u);
        x = x * 1664525u + 1013904223u;
        const int per = 1 + (int)((x >> 8) % 32u);
        x = x * 1664525u + 1013904223u;
        const int n = 1 + (int)((x >> 8) % 60000u);
        balanced(n, nctas, per);
        }
    strided(1000, 64);
    strided(5, 64);
    strided(4736, 4736);
    if (fails)
        {
        std::fprintf(stderr, "%d checks failed\n", fails);
        return 1;
        }
    std::printf("SLICE_ITER_OK\n");
    return 0;
    }
