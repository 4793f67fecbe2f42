// median5_stream.cuh -- exact 5x5 median of packed 8-bit images (3 or 4 bytes per pixel), replicate border:
// smooth_cuda type=median:kw=5:kh=5 (vf_smooth_nvcv.c:288-296 hands the window to CV-CUDA's MedianBlur).
//
// Same layout as median3_stream.cuh: a row is a stream of bytes (the horizontal neighbours of byte k are bytes
// k +- BPP, k +- 2 BPP), a thread owns S = NOUT * BPP consecutive byte columns of TWO row bands, and every
// register holds one byte of band A in its low u16 lane and the byte at the same column of band B in its high
// lane (each lane = 257 x byte), so each VIMNMX / VIMNMX3 ranks two samples.  The 25-sample selection is split
// so that most of it is shared between neighbouring outputs (tools/gen_median5_stream.py generates and
// verifies the min/max blocks, median5_nets.inc):
//   * every byte column of the 5 source rows is sorted once (9 compare-exchanges), shared by the 5 outputs
//     whose windows contain it;
//   * outputs come in horizontal pairs (x, x+1) of the same channel: their windows share 4 columns.  Those are
//     merged two by two into sorted 10s (26 instructions; each merge serves two pairs), and only ranks 7..12 of
//     the 20 shared samples can be the median of either window (a sample of rank r among the 20 has rank
//     r..r+5 among the 25), which a pruned Batcher merge extracts in sorted order (34 instructions per pair);
//   * per output: median = rank 5 of (6 candidates + its own sorted 5th column) = min over the six splits of
//     the larger prefix end (8 instructions).
// ~34-40 min/max instructions per output sample against 101 for the selection network it replaces
// (median_net_kernel<.,5,5>), no shared memory, no barriers; the five source rows of a step are re-read
// through L1 (the ALU pipe, not memory, bounds the kernel: VIMNMX issues at 64 lanes / clk / SM).
#pragma once
#include "common.cuh"
#include "median3_stream.cuh"

namespace gmatb {

#include "median5_nets.inc"

__device__ __forceinline__ void med5_cx(unsigned &a, unsigned &b) { const unsigned lo = mmin2(a, b), hi = mmax2(a, b); a = lo; b = hi; }
// 9 compare-exchanges
__device__ __forceinline__ void med5_sort5(unsigned (&v)[5]) {
    med5_cx(v[0], v[1]); med5_cx(v[3], v[4]); med5_cx(v[2], v[4]); med5_cx(v[2], v[3]); med5_cx(v[0], v[3]);
    med5_cx(v[0], v[2]); med5_cx(v[1], v[4]); med5_cx(v[1], v[3]); med5_cx(v[1], v[2]);
}

// one channel phase: NOUT + 4 unsorted columns (5 rows each) -> the NOUT medians of the columns 2 .. NOUT+1
template <int NOUT>
__device__ __forceinline__ void med5_phase(unsigned (&sc)[NOUT + 4][5], bool ledge, bool redge, unsigned (&med)[NOUT]) {
#pragma unroll
    for (int m = 0; m < NOUT + 4; m++) med5_sort5(sc[m]);
    if (ledge) {                                       // replicate border: the two columns left of the frame are column 2
#pragma unroll
        for (int r = 0; r < 5; r++) { sc[0][r] = sc[2][r]; sc[1][r] = sc[2][r]; }
    }
    if (redge) {
#pragma unroll
        for (int r = 0; r < 5; r++) { sc[NOUT + 2][r] = sc[NOUT + 1][r]; sc[NOUT + 3][r] = sc[NOUT + 1][r]; }
    }
    unsigned pm[NOUT / 2 + 1][10];
#pragma unroll
    for (int k = 0; k <= NOUT / 2; k++) med5_merge55(sc[2 * k + 1], sc[2 * k + 2], pm[k]);
#pragma unroll
    for (int t = 0; t < NOUT / 2; t++) {
        unsigned cand[6], ml[1], mr[1];
        med5_select6(pm[t], pm[t + 1], cand);
        med5_final(cand, sc[2 * t], ml);
        med5_final(cand, sc[2 * t + 5], mr);
        med[2 * t] = ml[0]; med[2 * t + 1] = mr[0];
    }
}

template <int BPP, int NOUT>
// resident CTAs per SM: 3-byte pixels 3 (168 registers, 64 bytes of spills: 130 -> 134 Gpx/s), 4-byte pixels 2 (3 measured 88 -> 80)
__global__ void __launch_bounds__(128, BPP == 3 ? 3 : 2) median5_stream_kernel(const Med3Params P) {
    static_assert(NOUT % 2 == 0 && (NOUT * BPP) % 4 == 0, "pairs of outputs, whole words");
    constexpr int S = NOUT * BPP;                      // output byte columns per thread
    constexpr int HB = 2 * BPP;                        // halo bytes per side
    constexpr int W0 = 8;                              // the strip's words start 8 bytes before c0 (covers 2 BPP <= 8)
    constexpr int NW = (W0 + S + HB + 3) / 4;          // aligned words covering [c0 - 8, c0 + S + 2 BPP)
    constexpr int WEND = (W0 + S) / 4;                 // first word past the strip's own bytes
    constexpr int NCOL = NOUT + 4;                     // columns of one channel phase
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * S;
    if (c0 >= P.wb) return;
    const long long fz = blockIdx.z;
    const int ya = blockIdx.y * 2 * P.rows, yb = ya + P.rows;      // first output row of band A / band B
    const uint8_t *ps = P.sp + fz * P.sbs + c0 - W0;
    uint8_t *pd = P.dp + fz * P.dbs + c0;
    const bool ledge = c0 == 0, redge = c0 + S == P.wb;
    const int H = P.H;

#pragma unroll 1
    for (int i = 0; i < P.rows; i++) {
        if (ya + i >= H) break;
        // raw words of the five window rows of both bands (clamped rows; words outside the row are never used as data)
        uint32_t wa[5][NW], wb_[5][NW];
#pragma unroll
        for (int r = 0; r < 5; r++) {
            const int ra = min(max(ya + i + r - 2, 0), H - 1), rb = min(max(yb + i + r - 2, 0), H - 1);
            const uint32_t *qa = reinterpret_cast<const uint32_t *>(ps + (size_t)ra * P.spitch);
            const uint32_t *qb = reinterpret_cast<const uint32_t *>(ps + (size_t)rb * P.spitch);
#pragma unroll
            for (int k = 0; k < NW; k++) {
                const int wi = (k < 2 && ledge) ? 2 : (k >= WEND && redge) ? WEND - 1 : k;
                wa[r][k] = __ldg(qa + wi); wb_[r][k] = __ldg(qb + wi);
            }
        }
        uint32_t oa[S / 4], ob[S / 4];
        if constexpr (BPP == 4) {
            // a word is a pixel: channel c is byte c of every word, so the channel phases differ only in their PRMT
            // selectors and run as a real loop (a quarter of the unrolled code: the step stays inside the instruction cache)
#pragma unroll
            for (int q = 0; q < S / 4; q++) { oa[q] = 0; ob[q] = 0; }
#pragma unroll 1
            for (int c = 0; c < 4; c++) {
                const unsigned sel_in = 0x4400u + c * 0x1111u;                  // {a_c, a_c, b_c, b_c}
                const unsigned ins_a = 0x3210u ^ ((unsigned)(c ^ 4) << (4 * c)), ins_b = 0x3210u ^ ((unsigned)(c ^ 6) << (4 * c));
                unsigned sc[NCOL][5], med[NOUT];
#pragma unroll
                for (int m = 0; m < NCOL; m++)
#pragma unroll
                    for (int r = 0; r < 5; r++) sc[m][r] = __byte_perm(wa[r][m], wb_[r][m], sel_in);
                med5_phase<NOUT>(sc, ledge, redge, med);
#pragma unroll
                for (int m = 0; m < NOUT; m++) { oa[m] = __byte_perm(oa[m], med[m], ins_a); ob[m] = __byte_perm(ob[m], med[m], ins_b); }
            }
        } else {
            unsigned o[S];
#pragma unroll
            for (int c = 0; c < BPP; c++) {            // channel phase: byte columns c0 + c + (m - 2) BPP, m = 0 .. NCOL-1
                unsigned sc[NCOL][5], med[NOUT];
#pragma unroll
                for (int m = 0; m < NCOL; m++) {
                    const int b = W0 - HB + c + m * BPP;   // byte offset from c0 - 8
#pragma unroll
                    for (int r = 0; r < 5; r++)
                        sc[m][r] = __byte_perm(wa[r][b >> 2], wb_[r][b >> 2], (b & 3) * 0x0011u + 0x4400u + (b & 3) * 0x1100u);
                }
                med5_phase<NOUT>(sc, ledge, redge, med);
#pragma unroll
                for (int m = 0; m < NOUT; m++) o[c + m * BPP] = med[m];
            }
#pragma unroll
            for (int q = 0; q < S / 4; q++) {
                const unsigned p01 = __byte_perm(o[4 * q], o[4 * q + 1], 0x6240u);      // {a0, a1, b0, b1}
                const unsigned p23 = __byte_perm(o[4 * q + 2], o[4 * q + 3], 0x6240u);
                oa[q] = __byte_perm(p01, p23, 0x5410u);
                ob[q] = __byte_perm(p01, p23, 0x7632u);
            }
        }
        uint8_t *da = pd + (size_t)(ya + i) * P.dpitch, *db = pd + (size_t)(yb + i) * P.dpitch;
        const bool wb2 = yb + i < H;
        if (S % 16 == 0) {
#pragma unroll
            for (int q = 0; q < S / 16; q++) {
                stg128(da + 16 * q, make_uint4(oa[4 * q], oa[4 * q + 1], oa[4 * q + 2], oa[4 * q + 3]));
                if (wb2) stg128(db + 16 * q, make_uint4(ob[4 * q], ob[4 * q + 1], ob[4 * q + 2], ob[4 * q + 3]));
            }
        } else if (S % 8 == 0) {
#pragma unroll
            for (int q = 0; q < S / 8; q++) {
                stg64(da + 8 * q, make_uint2(oa[2 * q], oa[2 * q + 1]));
                if (wb2) stg64(db + 8 * q, make_uint2(ob[2 * q], ob[2 * q + 1]));
            }
        } else {
#pragma unroll
            for (int q = 0; q < S / 4; q++) {
                stg32(da + 4 * q, oa[q]);
                if (wb2) stg32(db + 4 * q, ob[q]);
            }
        }
    }
}

}  // namespace gmatb
