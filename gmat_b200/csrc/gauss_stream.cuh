// gauss_stream.cuh -- separable Gaussian for the common small kernels (3/5/7 taps per axis),
// register-streaming like the fused scale kernel: a thread owns 4 pixels of a row strip and
// walks down a band of rows.
//   * per source row: (4 + KW - 1) pixels are fetched as aligned 32-bit words (neighbouring
//     threads overlap -> L1 hits) ONE ROW AHEAD of their use, bytes go to float through PRMT
//     magic numbers + one packed FADD2, the KW-tap horizontal chain runs packed on component pairs;
//   * vertically every output row has a running accumulator; a source row updates the KH
//     accumulators it contributes to, in the oracle's tap order (the rows of an output arrive
//     in increasing tap index), the one that completes is rounded (FADD magic, ties-to-even =
//     rintf), packed and stored.  No shared memory, no barriers.
//   * only strips whose whole window lies inside the row run here (thread index t0 <= t < t1); the
//     few frame-edge columns, unaligned images and any other kernel size use gaussian_kernel
//     (filters.cu), so this kernel carries no horizontal border code at all.  Rows above / below
//     the frame map through border_idx (out-of-line, executed only by the first / last band).
// Arithmetic is exactly oracle/gmat_oracle.c orc_gaussian: acc = fmaf(k[i], v, acc) from 0.0f,
// horizontal then vertical, rintf, saturate.
#pragma once
#include "common.cuh"

namespace gmatb {

struct GaussS { float kx[7], ky[7]; int border; int band; };

__device__ __noinline__ int border_row(int y, int H, int mode) { return border_idx(y, H, mode); }

// INTERIOR: every source row of the band lies inside the frame, so the row fetch is straight-line code (no border
// mapping, no zero rows) and ptxas issues the loads at the top of a row's step, a whole step ahead of their use.
// With the border branch in front of them it sank the loads to ~50 instructions before their consumer
// (ncu: 43 % of all stall samples on that consumer, 33 % issue-active).
// PF2: loads are issued two rows ahead of their use (three row buffers) instead of one.  Measured on B200, 4K, 5x5:
// rgb24 (96 registers, 5 CTAs/SM) 433 Gpx/s with one row ahead, 397 with two (126 registers, 4 CTAs/SM);
// rgba 289 -> 341 Gpx/s: the 4-byte pixel form is the latency-bound one.
// RING > 0: the rows are fetched RING - 1 steps ahead by 4-byte cp.async copies into a per-thread column of a shared-memory
// ring (slot-major, [slot][word][thread]: conflict-free, and a thread only ever reads what it copied itself, so
// cp.async.wait_group is the only synchronisation: no barrier, no mbarrier) and picked up by LDS when their step
// comes.  The register forms above can keep one or two rows in flight (one step is ~150 instructions; 4-5 resident
// warps per scheduler cover ~750 issue cycles, about one loaded-DRAM latency: ncu long-scoreboard 3.2 per issue);
// the ring keeps RING - 1 rows in flight at no register cost.
template <int BPP, int KW, int KH, bool INTERIOR, bool PF2, int RING = 0>
__device__ __forceinline__ void gauss_stream_band(const uint8_t *sp, int spitch, long long sbs,
                                                  uint8_t *dp, int dpitch, long long dbs, int H, int t0, int t1, const GaussS &G) {
    constexpr int RX = KW / 2, RY = KH / 2, NPX = 4;
    constexpr int NWIN = (NPX + KW - 1) * BPP;                   // window components per row
    constexpr int NOUT = NPX * BPP;                              // output components per row (12 or 16)
    constexpr int WOFF = RX * BPP;                               // bytes before the strip
    constexpr int WSH = (4 - (WOFF & 3)) & 3;                    // window starts WSH bytes into the first aligned word
    constexpr int NWORDS = (WSH + NWIN + 3) / 4;
    const int t = t0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= t1) return;
    const int x0 = t * NPX;
    const long long fz = blockIdx.z;
    const int y_begin = blockIdx.y * G.band, y_end = min(y_begin + G.band, H);
    const uint8_t *ps = sp + fz * sbs + (size_t)x0 * BPP - WOFF - WSH;      // word aligned (host checks)
    uint8_t *pd = dp + fz * dbs + (size_t)x0 * BPP;

    f2 acc[KH][NOUT / 2];
#pragma unroll
    for (int a = 0; a < KH; a++)
#pragma unroll
        for (int o = 0; o < NOUT / 2; o++) acc[a][o] = 0ull;

    const int nrows = (y_end - y_begin) + KH - 1;
    uint32_t w[NWORDS], wn[NWORDS], wnn[NWORDS];   // rows i, i+1, i+2: loads are issued two steps ahead of their use
    extern __shared__ uint32_t gs_ring[];
    const uint32_t ring0 = (uint32_t)__cvta_generic_to_shared(gs_ring) + threadIdx.x * 4u;
    // RING: copy row i into slot i % RING (one commit group per call, also when nothing is copied)
    auto fetch_ring = [&](int i) {
        int sy = y_begin - RY + i;
        const uint32_t sa = ring0 + (uint32_t)((i % (RING > 0 ? RING : 1)) * NWORDS) * 512u;
        bool zero = false;
        if (!INTERIOR) {
            if ((unsigned)sy >= (unsigned)H) sy = border_row(sy, H, G.border);
            zero = sy < 0;
        } else sy = min(sy, H - 1);
        if (i < nrows) {
            if (zero) {
#pragma unroll
                for (int k = 0; k < NWORDS; k++) asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa + k * 512u), "r"(0u) : "memory");
            } else {
                const uint8_t *q = ps + (size_t)sy * spitch;
#pragma unroll
                for (int k = 0; k < NWORDS; k++)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + k * 512u), "l"(q + 4 * k) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto take_ring = [&](int i, uint32_t (&dst)[NWORDS]) {
        asm volatile("cp.async.wait_group %0;" ::"n"(RING > 1 ? RING - 2 : 0) : "memory");
        const uint32_t sa = ring0 + (uint32_t)((i % (RING > 0 ? RING : 1)) * NWORDS) * 512u;
#pragma unroll
        for (int k = 0; k < NWORDS; k++) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(dst[k]) : "r"(sa + k * 512u) : "memory");
    };
    auto fetch = [&](int i, uint32_t (&dst)[NWORDS]) {
        int sy = y_begin - RY + i;
        if (INTERIOR) {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(ps + (size_t)min(sy, H - 1) * spitch);
#pragma unroll
            for (int k = 0; k < NWORDS; k++) dst[k] = __ldg(q + k);
            // one step (~150 instructions x 4 resident warps) is about one loaded-DRAM latency: pull the row three
            // steps ahead into L2 so that the register prefetch above becomes an L2 hit
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ps + (size_t)min(sy + 3, H - 1) * spitch + WSH));
            return;
        }
        if ((unsigned)sy >= (unsigned)H) sy = border_row(sy, H, G.border);
        if (sy < 0 || i >= nrows) {                              // BORDER_CONSTANT row (or past the band): zeros
#pragma unroll
            for (int k = 0; k < NWORDS; k++) dst[k] = 0u;
        } else {
            const uint32_t *q = reinterpret_cast<const uint32_t *>(ps + (size_t)sy * spitch);
#pragma unroll
            for (int k = 0; k < NWORDS; k++) dst[k] = __ldg(q + k);
        }
    };
    if (RING > 0) {
#pragma unroll 1
        for (int i = 0; i < RING - 1; i++) fetch_ring(i);
    } else {
        fetch(0, w);
        if (PF2) fetch(1, wn);
    }
    for (int i0 = 0; i0 < nrows; i0 += KH) {
#pragma unroll
        for (int ph = 0; ph < KH; ph++) {
            const int i = i0 + ph;
            if (i >= nrows) break;
            if (RING > 0) { take_ring(i, w); fetch_ring(i + RING - 1); }   // into the slot read one step ago
            else if (PF2) fetch(i + 2, wnn); else fetch(i + 1, wn);
            // ---- window of this row as floats: even-aligned pairs E, odd-aligned pairs O (BPP 3 needs both) ----
            f2 E[(NWIN + 1) / 2], O[(NWIN + 1) / 2];
            {
                float m[NWIN + 1];
#pragma unroll
                for (int c = 0; c < NWIN; c++) {
                    const int b = WSH + c;
                    m[c] = (b & 3) == 0 ? byte_magic<0>(w[b >> 2]) : (b & 3) == 1 ? byte_magic<1>(w[b >> 2])
                         : (b & 3) == 2 ? byte_magic<2>(w[b >> 2]) : byte_magic<3>(w[b >> 2]);
                }
                m[NWIN] = GMATB_MAGIC;
#pragma unroll
                for (int j = 0; j < (NWIN + 1) / 2; j++) E[j] = add2(pk(m[2 * j], m[2 * j + 1]), bc(-GMATB_MAGIC));
                if (BPP & 1) {
#pragma unroll
                    for (int j = 0; j < NWIN / 2; j++) O[j] = add2(pk(m[2 * j + 1], m[2 * j + 2]), bc(-GMATB_MAGIC));
                }
            }
            // ---- horizontal chain on output component pairs ------------------------------------------------
            f2 T[NOUT / 2];
#pragma unroll
            for (int o = 0; o < NOUT / 2; o++) {
                f2 a = 0ull;
#pragma unroll
                for (int k = 0; k < KW; k++) {
                    const int c = 2 * o + k * BPP;                 // first window component of the pair
                    const f2 v = (c & 1) ? O[c >> 1] : E[c >> 1];
                    a = fma2(bc(G.kx[k]), v, a);
                }
                T[o] = a;
            }
            // ---- vertical running accumulators ---------------------------------------------------------------
#pragma unroll
            for (int k = 0; k < KH; k++) {
                const int slot = (ph - k + KH) % KH;
#pragma unroll
                for (int o = 0; o < NOUT / 2; o++) acc[slot][o] = fma2(bc(G.ky[k]), T[o], k == 0 ? (f2)0ull : acc[slot][o]);
            }
            // the accumulator that just received tap KH-1 is complete: output row y = y_begin + i - (KH-1)
            const int yo = y_begin + i - (KH - 1);
            if (yo >= y_begin) {
                const int sl = (ph - (KH - 1) + KH) % KH;
                uint32_t r[NOUT];
#pragma unroll
                for (int o = 0; o < NOUT / 2; o++) {
                    int b0, b1;
                    upki(add2(acc[sl][o], bc(GMATB_MAGIC)), b0, b1);          // 2^23 + rint(v): low byte = pixel (v in [0,255.0x])
                    r[2 * o] = (uint32_t)b0; r[2 * o + 1] = (uint32_t)b1;
                }
                uint32_t *q = reinterpret_cast<uint32_t *>(pd + (size_t)yo * dpitch);
#pragma unroll
                for (int wv = 0; wv < NOUT / 4; wv++) {
                    uint32_t lo, hi, word;
                    asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(lo) : "r"(r[4 * wv]), "r"(r[4 * wv + 1]));
                    asm("prmt.b32 %0, %1, %2, 0x0040;" : "=r"(hi) : "r"(r[4 * wv + 2]), "r"(r[4 * wv + 3]));
                    asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(word) : "r"(lo), "r"(hi));
                    q[wv] = word;
                }
            }
            if (RING == 0) {
#pragma unroll
                for (int k = 0; k < NWORDS; k++) { w[k] = wn[k]; if (PF2) wn[k] = wnn[k]; }
            }
        }
    }
}

// words of one ring slot per thread: the host sizes the dynamic shared memory with it
constexpr int gauss_ring_words(int bpp, int kw) {
    return ((4 - ((kw / 2 * bpp) & 3)) & 3) + (4 + kw - 1) * bpp + 3 >> 2;
}

template <int BPP, int KW, int KH, int RING>
__global__ void __launch_bounds__(128, BPP == 4 ? 4 : 5) gauss_ring_kernel(const uint8_t *sp, int spitch, long long sbs,
                                                             uint8_t *dp, int dpitch, long long dbs, int H, int t0, int t1, GaussS G) {
    const int y_begin = blockIdx.y * G.band, y_end = min(y_begin + G.band, H);
    if (y_begin - KH / 2 >= 0 && y_end + KH / 2 <= H) gauss_stream_band<BPP, KW, KH, true, false, RING>(sp, spitch, sbs, dp, dpitch, dbs, H, t0, t1, G);
    else                                             gauss_stream_band<BPP, KW, KH, false, false, RING>(sp, spitch, sbs, dp, dpitch, dbs, H, t0, t1, G);
}

template <int BPP, int KW, int KH, int MINB>
__global__ void __launch_bounds__(128, MINB) gauss_stream_kernel(const uint8_t *sp, int spitch, long long sbs,
                                                               uint8_t *dp, int dpitch, long long dbs, int H, int t0, int t1, GaussS G) {
    const int y_begin = blockIdx.y * G.band, y_end = min(y_begin + G.band, H);
    if (y_begin - KH / 2 >= 0 && y_end + KH / 2 <= H) gauss_stream_band<BPP, KW, KH, true, MINB == 4>(sp, spitch, sbs, dp, dpitch, dbs, H, t0, t1, G);
    else                                             gauss_stream_band<BPP, KW, KH, false, MINB == 4>(sp, spitch, sbs, dp, dpitch, dbs, H, t0, t1, G);
}

}  // namespace gmatb
