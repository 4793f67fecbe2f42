// median3_stream.cuh -- exact 3x3 median of packed 8-bit images (3 or 4 bytes per pixel), replicate
// border: the smooth filter's default window, and what the reference's smooth_nvcv always ends up
// running (its type switch falls through to MedianBlur, vf_smooth_nvcv.c:130-138,288-296).
//
// A row is treated as a stream of bytes: the horizontal neighbours of byte k are bytes k-BPP and k+BPP,
// so pixels and channels disappear.  A thread owns 16 consecutive byte columns of TWO row bands at once:
// every register holds one byte of band A in its low half and the byte at the same column of band B in
// its high half (u16x2, each lane = 257 x byte), so each VIMNMX / VIMNMX3 compares two samples and the horizontal shifts never
// cross halves.  Walking down, per source row: 16 + 2*BPP byte columns are loaded as aligned words and
// expanded (one PRMT per column pair); the column is sorted with the two rows above (lo/mid/hi: 6 ops,
// shared by the three outputs that use the column); per output byte
//      median = med3( max3(lo[-],lo[0],lo[+]),  med3(mid[-],mid[0],mid[+]),  min3(hi[-],hi[0],hi[+]) )
// (10 ops with the 3-input forms).  No shared memory, no barriers; ~11 instructions per output byte
// (the shared-memory selection-network kernel it replaces for 3x3 needed ~37).
#pragma once
#include "common.cuh"

namespace gmatb {

__device__ __forceinline__ unsigned mmin2(unsigned a, unsigned b) { return __vminu2(a, b); }
__device__ __forceinline__ unsigned mmax2(unsigned a, unsigned b) { return __vmaxu2(a, b); }
__device__ __forceinline__ unsigned mmin3(unsigned a, unsigned b, unsigned c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned mmax3(unsigned a, unsigned b, unsigned c) { return __vimax3_u16x2(a, b, c); }
__device__ __forceinline__ unsigned mmed3(unsigned a, unsigned b, unsigned c) {
    return mmax3(mmin2(a, b), mmin2(b, c), mmin2(a, c));
}

struct Med3Params {
    const uint8_t *sp; uint8_t *dp;
    int spitch, dpitch;
    long long sbs, dbs;
    int wb;          // row length in bytes (multiple of 16)
    int H;
    int rows;        // rows per band; a CTA row covers 2 * rows output rows
};

template <int BPP>
__global__ void __launch_bounds__(128, 4) median3_stream_kernel(const Med3Params P) {
    constexpr int S = 16, NC = S + 2 * BPP;          // output byte columns per thread / columns incl. halo
    constexpr int W0 = 4;                             // the strip's words start 4 bytes before c0 (covers BPP <= 4)
    constexpr int NW = (W0 + S + BPP + 3) / 4;        // aligned words covering [c0-4, c0+S+BPP)
    const int c0 = (blockIdx.x * blockDim.x + threadIdx.x) * S;
    if (c0 >= P.wb) return;
    const long long fz = blockIdx.z;
    const int ya = blockIdx.y * 2 * P.rows, yb = ya + P.rows;      // first output row of band A / band B
    const uint8_t *ps = P.sp + fz * P.sbs;
    uint8_t *pd = P.dp + fz * P.dbs + c0;
    const bool ledge = c0 == 0, redge = c0 + S == P.wb;
    const int H = P.H;

    // one source row of both bands -> NC registers, byte of band A | byte of band B << 16
    auto load_row = [&](int ra, int rb, unsigned (&e)[NC]) {
        ra = min(max(ra, 0), H - 1); rb = min(max(rb, 0), H - 1);
        const uint32_t *qa = reinterpret_cast<const uint32_t *>(ps + (size_t)ra * P.spitch + c0);
        const uint32_t *qb = reinterpret_cast<const uint32_t *>(ps + (size_t)rb * P.spitch + c0);
        uint32_t wa[NW], wb_[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) {
            // words before the row start / past its end are never used as data (edge columns are replicated below):
            // read a word that exists instead
            const int wi = (k == 0 && ledge) ? 0 : (k == NW - 1 && redge) ? NW - 3 : k - 1;
            wa[k] = __ldg(qa + wi); wb_[k] = __ldg(qb + wi);
        }
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const int b = W0 - BPP + j;               // byte offset from c0 - 4
            // {a, a, b, b}: each u16 lane is 257 * byte -- the same order as the bytes, and no zero byte is needed
            e[j] = __byte_perm(wa[b >> 2], wb_[b >> 2], (b & 3) * 0x0011u + 0x4400u + (b & 3) * 0x1100u);
        }
        if (ledge) {
#pragma unroll
            for (int j = 0; j < BPP; j++) e[j] = e[j + BPP];
        }
        if (redge) {
#pragma unroll
            for (int j = 0; j < BPP; j++) e[S + BPP + j] = e[S + j];
        }
    };

    // one output row of each band: rows (a, b) are the two above, c receives the row below
    auto step = [&](const unsigned (&a)[NC], const unsigned (&b)[NC], unsigned (&c)[NC], int i) {
        load_row(ya + i + 1, yb + i + 1, c);
        unsigned lo[NC], mid[NC], hi[NC];
#pragma unroll
        for (int j = 0; j < NC; j++) {
            const unsigned mn = mmin2(a[j], b[j]), mx = mmax2(a[j], b[j]);
            lo[j] = mmin2(mn, c[j]);
            const unsigned t = mmax2(mn, c[j]);
            mid[j] = mmin2(mx, t); hi[j] = mmax2(mx, t);
        }
        unsigned o[S];
#pragma unroll
        for (int k = 0; k < S; k++) {
            const unsigned l3 = mmax3(lo[k], lo[k + BPP], lo[k + 2 * BPP]);
            const unsigned h3 = mmin3(hi[k], hi[k + BPP], hi[k + 2 * BPP]);
            const unsigned m3 = mmed3(mid[k], mid[k + BPP], mid[k + 2 * BPP]);
            o[k] = mmed3(l3, m3, h3);
        }
        uint32_t oa[4], ob[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const unsigned p01 = __byte_perm(o[4 * q], o[4 * q + 1], 0x6240u);      // {a0, a1, b0, b1}
            const unsigned p23 = __byte_perm(o[4 * q + 2], o[4 * q + 3], 0x6240u);
            oa[q] = __byte_perm(p01, p23, 0x5410u);
            ob[q] = __byte_perm(p01, p23, 0x7632u);
        }
        if (ya + i < H) stg128(pd + (size_t)(ya + i) * P.dpitch, make_uint4(oa[0], oa[1], oa[2], oa[3]));
        if (yb + i < H) stg128(pd + (size_t)(yb + i) * P.dpitch, make_uint4(ob[0], ob[1], ob[2], ob[3]));
    };
    unsigned r0[NC], r1[NC], r2[NC];
    load_row(ya - 1, yb - 1, r0);
    load_row(ya, yb, r1);
#pragma unroll 1
    for (int i = 0; i < P.rows; i += 3) {          // the three row buffers rotate by name: no copies
        step(r0, r1, r2, i);
        if (i + 1 < P.rows) step(r1, r2, r0, i + 1);
        if (i + 2 < P.rows) step(r2, r0, r1, i + 2);
    }
}

}  // namespace gmatb
