// csc.cu -- unscaled converters: yuv->rgb, rgb->yuv, yuv->yuv, rgb24<->bgr24.
//
// Replaces libswscale/cuda/{yuv2rgb_cuda,yuv2yuv_cuda,rgb2rgb_cuda_kernel}.cu of
// the reference.  Where the reference gives each thread a 2x2 pixel block with
// byte-granular LDG/STG (6 loads + 27 one-byte stores per thread for NV12->RGB24),
// every kernel here gives a thread an 8-pixel x 2-row tile: 64/128-bit coalesced
// row loads, packed FFMA2 arithmetic, saturating I2IP byte packs and 64/128-bit
// stores, one launch per (batch of) frame(s) with the frame index on blockIdx.z.
// All kernels are pure streaming kernels bounded by HBM bandwidth.
#include "csc_core.cuh"

namespace gmatb {


// ---------------------------------------------------------------------------
// tile loads: luma/chroma samples as "magic" floats (2^23 + sample)
// ---------------------------------------------------------------------------
template <int BITS> __device__ __forceinline__ float magic_of(unsigned v) {
    return __uint_as_float(0x4B000000u | v);
}

template <int L, int BITS>
__device__ __forceinline__ void load_yuv_tile(const Img &s, long long fz, int x0, int y0, bool vec,
                                              float (&ym)[2][8], float (&um)[4], float (&vm)[4]) {
    const int W = s.w, H = s.h;
    const int cw = (W + 1) >> 1, ch = (H + 1) >> 1;
    const uint8_t *py = s.pl[0].p + fz * s.pl[0].bstride;
    if (vec) {
        if (BITS == 8) {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                uint2 w = ldg64(py + (size_t)(y0 + r) * s.pl[0].pitch + x0);
                ym[r][0] = byte_magic<0>(w.x); ym[r][1] = byte_magic<1>(w.x);
                ym[r][2] = byte_magic<2>(w.x); ym[r][3] = byte_magic<3>(w.x);
                ym[r][4] = byte_magic<0>(w.y); ym[r][5] = byte_magic<1>(w.y);
                ym[r][6] = byte_magic<2>(w.y); ym[r][7] = byte_magic<3>(w.y);
            }
            if (L == L_NV12) {
                const uint8_t *pc = s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + x0;
                uint2 w = ldg64(pc);
                um[0] = byte_magic<0>(w.x); vm[0] = byte_magic<1>(w.x);
                um[1] = byte_magic<2>(w.x); vm[1] = byte_magic<3>(w.x);
                um[2] = byte_magic<0>(w.y); vm[2] = byte_magic<1>(w.y);
                um[3] = byte_magic<2>(w.y); vm[3] = byte_magic<3>(w.y);
            } else {
                uint32_t wu = ldg32(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + (x0 >> 1));
                uint32_t wv = ldg32(s.pl[2].p + fz * s.pl[2].bstride + (size_t)(y0 >> 1) * s.pl[2].pitch + (x0 >> 1));
                um[0] = byte_magic<0>(wu); um[1] = byte_magic<1>(wu); um[2] = byte_magic<2>(wu); um[3] = byte_magic<3>(wu);
                vm[0] = byte_magic<0>(wv); vm[1] = byte_magic<1>(wv); vm[2] = byte_magic<2>(wv); vm[3] = byte_magic<3>(wv);
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; r++) {
                uint4 w = ldg128(py + (size_t)(y0 + r) * s.pl[0].pitch + x0 * 2);
                ym[r][0] = half_magic<0>(w.x); ym[r][1] = half_magic<1>(w.x);
                ym[r][2] = half_magic<0>(w.y); ym[r][3] = half_magic<1>(w.y);
                ym[r][4] = half_magic<0>(w.z); ym[r][5] = half_magic<1>(w.z);
                ym[r][6] = half_magic<0>(w.w); ym[r][7] = half_magic<1>(w.w);
            }
            if (L == L_NV12) {
                uint4 w = ldg128(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + x0 * 2);
                um[0] = half_magic<0>(w.x); vm[0] = half_magic<1>(w.x);
                um[1] = half_magic<0>(w.y); vm[1] = half_magic<1>(w.y);
                um[2] = half_magic<0>(w.z); vm[2] = half_magic<1>(w.z);
                um[3] = half_magic<0>(w.w); vm[3] = half_magic<1>(w.w);
            } else {
                uint2 wu = ldg64(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + x0);
                uint2 wv = ldg64(s.pl[2].p + fz * s.pl[2].bstride + (size_t)(y0 >> 1) * s.pl[2].pitch + x0);
                um[0] = half_magic<0>(wu.x); um[1] = half_magic<1>(wu.x); um[2] = half_magic<0>(wu.y); um[3] = half_magic<1>(wu.y);
                vm[0] = half_magic<0>(wv.x); vm[1] = half_magic<1>(wv.x); vm[2] = half_magic<0>(wv.y); vm[3] = half_magic<1>(wv.y);
            }
        }
        return;
    }
    // edge / unaligned tiles: clamped scalar loads
    constexpr int BS = BITS / 8;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        int yy = min(y0 + r, H - 1);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int xx = min(x0 + i, W - 1);
            const uint8_t *q = py + (size_t)yy * s.pl[0].pitch + xx * BS;
            unsigned v = BITS == 8 ? *q : *reinterpret_cast<const uint16_t *>(q);
            ym[r][i] = magic_of<BITS>(v);
        }
    }
    int cy = min(y0 >> 1, ch - 1);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int cx = min((x0 >> 1) + j, cw - 1);
        unsigned u, v;
        if (L == L_NV12) {
            const uint8_t *q = s.pl[1].p + fz * s.pl[1].bstride + (size_t)cy * s.pl[1].pitch + cx * 2 * BS;
            if (BITS == 8) { u = q[0]; v = q[1]; }
            else { u = reinterpret_cast<const uint16_t *>(q)[0]; v = reinterpret_cast<const uint16_t *>(q)[1]; }
        } else {
            const uint8_t *qu = s.pl[1].p + fz * s.pl[1].bstride + (size_t)cy * s.pl[1].pitch + cx * BS;
            const uint8_t *qv = s.pl[2].p + fz * s.pl[2].bstride + (size_t)cy * s.pl[2].pitch + cx * BS;
            if (BITS == 8) { u = *qu; v = *qv; }
            else { u = *reinterpret_cast<const uint16_t *>(qu); v = *reinterpret_cast<const uint16_t *>(qv); }
        }
        um[j] = magic_of<BITS>(u); vm[j] = magic_of<BITS>(v);
    }
}

// ---------------------------------------------------------------------------
// packed-RGB row stores (8 pixels), r/g/b = unclamped truncated integers
// ---------------------------------------------------------------------------
template <int DST, int SBITS>
__device__ __forceinline__ void store_rgb_row8(uint8_t *p, const int (&r)[8], const int (&g)[8], const int (&b)[8]) {
    constexpr bool SW = dst_swap(DST);
    int c0[8], c1[8], c2[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int rr = r[i], gg = g[i], bb = b[i];
        if (SBITS == 16 && !dst_is16(DST)) { rr = clamp_i(rr, 65535) >> 8; gg = clamp_i(gg, 65535) >> 8; bb = clamp_i(bb, 65535) >> 8; }
        c0[i] = SW ? bb : rr; c1[i] = gg; c2[i] = SW ? rr : bb;
    }
    if (DST == D_RGB24 || DST == D_BGR24 || ((DST == D_RGB48 || DST == D_BGR48) && SBITS == 8)) {
        uint32_t w[6];
        w[0] = pack4_u8(c0[0], c1[0], c2[0], c0[1]);
        w[1] = pack4_u8(c1[1], c2[1], c0[2], c1[2]);
        w[2] = pack4_u8(c2[2], c0[3], c1[3], c2[3]);
        w[3] = pack4_u8(c0[4], c1[4], c2[4], c0[5]);
        w[4] = pack4_u8(c1[5], c2[5], c0[6], c1[6]);
        w[5] = pack4_u8(c2[6], c0[7], c1[7], c2[7]);
        if (DST == D_RGB24 || DST == D_BGR24) {
            stg64(p, make_uint2(w[0], w[1])); stg64(p + 8, make_uint2(w[2], w[3])); stg64(p + 16, make_uint2(w[4], w[5]));
        } else {   // 8-bit source -> 16-bit components: c << 8 (yuv2rgb_cuda.cu:95-99)
            uint32_t e[12];
#pragma unroll
            for (int k = 0; k < 6; k++) { e[2 * k] = prmt(w[k], 0xFFu, 0x1505u); e[2 * k + 1] = prmt(w[k], 0xFFu, 0x3525u); }
            stg128(p, make_uint4(e[0], e[1], e[2], e[3])); stg128(p + 16, make_uint4(e[4], e[5], e[6], e[7]));
            stg128(p + 32, make_uint4(e[8], e[9], e[10], e[11]));
        }
    } else if (DST == D_RGBA || DST == D_BGRA) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = pack4_u8(c0[i], c1[i], c2[i], 255);
        stg128(p, make_uint4(w[0], w[1], w[2], w[3])); stg128(p + 16, make_uint4(w[4], w[5], w[6], w[7]));
    } else if (DST == D_RGB48 || DST == D_BGR48) {   // 16-bit source
        uint32_t e[12];
        e[0] = pack2_u16(c0[0], c1[0]); e[1] = pack2_u16(c2[0], c0[1]); e[2] = pack2_u16(c1[1], c2[1]);
        e[3] = pack2_u16(c0[2], c1[2]); e[4] = pack2_u16(c2[2], c0[3]); e[5] = pack2_u16(c1[3], c2[3]);
        e[6] = pack2_u16(c0[4], c1[4]); e[7] = pack2_u16(c2[4], c0[5]); e[8] = pack2_u16(c1[5], c2[5]);
        e[9] = pack2_u16(c0[6], c1[6]); e[10] = pack2_u16(c2[6], c0[7]); e[11] = pack2_u16(c1[7], c2[7]);
        stg128(p, make_uint4(e[0], e[1], e[2], e[3])); stg128(p + 16, make_uint4(e[4], e[5], e[6], e[7]));
        stg128(p + 32, make_uint4(e[8], e[9], e[10], e[11]));
    } else {   // RGBA64 / BGRA64; alpha is 255 (not 65535): reference quirk, yuv2rgb_cuda.cu:89
#pragma unroll
        for (int h = 0; h < 4; h++) {
            uint32_t e[4];
#pragma unroll
            for (int k = 0; k < 2; k++) {
                int i = 2 * h + k;
                if (SBITS == 8) {
                    uint32_t w = pack4_u8(c0[i], c1[i], c2[i], 0);
                    e[2 * k] = prmt(w, 0xFFu, 0x1505u);       // {0,c0,0,c1}
                    e[2 * k + 1] = prmt(w, 0xFFu, 0x5425u);   // {0,c2,0xFF,0}
                } else {
                    e[2 * k] = pack2_u16(c0[i], c1[i]);
                    e[2 * k + 1] = pack2_u16(c2[i], 255);
                }
            }
            stg128(p + 16 * h, make_uint4(e[0], e[1], e[2], e[3]));
        }
    }
}

template <int DST, int SBITS>
__device__ __forceinline__ void store_rgb_px(uint8_t *p, int r, int g, int b) {
    constexpr int SMAX = SBITS == 8 ? 255 : 65535;
    r = clamp_i(r, SMAX); g = clamp_i(g, SMAX); b = clamp_i(b, SMAX);
    if (dst_is16(DST)) {
        if (SBITS == 8) { r <<= 8; g <<= 8; b <<= 8; }
        uint16_t *q = reinterpret_cast<uint16_t *>(p);
        q[0] = dst_swap(DST) ? b : r; q[1] = g; q[2] = dst_swap(DST) ? r : b;
        if (dst_alpha(DST)) q[3] = 255;
    } else {
        if (SBITS == 16) { r >>= 8; g >>= 8; b >>= 8; }
        p[0] = dst_swap(DST) ? b : r; p[1] = g; p[2] = dst_swap(DST) ? r : b;
        if (dst_alpha(DST)) p[3] = 255;
    }
}

// ---------------------------------------------------------------------------
// yuv -> packed rgb
// ---------------------------------------------------------------------------
// FMAF: which of the reference's two roundings of the chain (csc_core.cuh): libgpuscale's 8-bit kernels compile to
// the FADD form, its 16-bit kernels and every kernel of metrans' NvCodec/ColorSpace.cu to the FMA form.
// raw words of one aligned 8 x 2 tile of an 8-bit source (the vector path of load_yuv_tile, split so that the loads of several
// tiles can be in flight while only 6 registers per tile are held)
struct RawTile8 { uint2 y0, y1, c; };
template <int L>
__device__ __forceinline__ RawTile8 load_raw_tile8(const Img &s, long long fz, int x0, int y0) {
    RawTile8 R;
    const uint8_t *py = s.pl[0].p + fz * s.pl[0].bstride + (size_t)y0 * s.pl[0].pitch + x0;
    R.y0 = ldg64(py); R.y1 = ldg64(py + s.pl[0].pitch);
    if (L == L_NV12) R.c = ldg64(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + x0);
    else {
        R.c.x = ldg32(s.pl[1].p + fz * s.pl[1].bstride + (size_t)(y0 >> 1) * s.pl[1].pitch + (x0 >> 1));
        R.c.y = ldg32(s.pl[2].p + fz * s.pl[2].bstride + (size_t)(y0 >> 1) * s.pl[2].pitch + (x0 >> 1));
    }
    return R;
}
template <int L>
__device__ __forceinline__ void unpack_raw_tile8(const RawTile8 &R, float (&ym)[2][8], float (&um)[4], float (&vm)[4]) {
    ym[0][0] = byte_magic<0>(R.y0.x); ym[0][1] = byte_magic<1>(R.y0.x); ym[0][2] = byte_magic<2>(R.y0.x); ym[0][3] = byte_magic<3>(R.y0.x);
    ym[0][4] = byte_magic<0>(R.y0.y); ym[0][5] = byte_magic<1>(R.y0.y); ym[0][6] = byte_magic<2>(R.y0.y); ym[0][7] = byte_magic<3>(R.y0.y);
    ym[1][0] = byte_magic<0>(R.y1.x); ym[1][1] = byte_magic<1>(R.y1.x); ym[1][2] = byte_magic<2>(R.y1.x); ym[1][3] = byte_magic<3>(R.y1.x);
    ym[1][4] = byte_magic<0>(R.y1.y); ym[1][5] = byte_magic<1>(R.y1.y); ym[1][6] = byte_magic<2>(R.y1.y); ym[1][7] = byte_magic<3>(R.y1.y);
    if (L == L_NV12) {
        um[0] = byte_magic<0>(R.c.x); vm[0] = byte_magic<1>(R.c.x); um[1] = byte_magic<2>(R.c.x); vm[1] = byte_magic<3>(R.c.x);
        um[2] = byte_magic<0>(R.c.y); vm[2] = byte_magic<1>(R.c.y); um[3] = byte_magic<2>(R.c.y); vm[3] = byte_magic<3>(R.c.y);
    } else {
        um[0] = byte_magic<0>(R.c.x); um[1] = byte_magic<1>(R.c.x); um[2] = byte_magic<2>(R.c.x); um[3] = byte_magic<3>(R.c.x);
        vm[0] = byte_magic<0>(R.c.y); vm[1] = byte_magic<1>(R.c.y); vm[2] = byte_magic<2>(R.c.y); vm[3] = byte_magic<3>(R.c.y);
    }
}

// TILES: row pairs per thread.  The loads of every pair are issued before the first conversion: with one pair a thread has 24
// bytes in flight (18 KB per SM at 24 resident warps, half of what 6.5 TB/s x the loaded DRAM latency asks for; ncu:
// long-scoreboard 6.7 per issue, 41 % issue-active); 8-bit sources run two.
template <int L, int SBITS, int DST, bool SPARSE, bool FMAF = (SBITS == 16), int TILES = 1>
#ifndef GMATB_Y2R_MINB
#define GMATB_Y2R_MINB 4
#endif
__global__ void __launch_bounds__(256, TILES == 1 ? GMATB_Y2R_MINB : 3) yuv2rgb_kernel(Img src, Img dst, Mat9 M, int vec_ok) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int yb = (blockIdx.y * 8 + threadIdx.y) * 2 * TILES;
    if (x0 >= src.w || yb >= src.h) return;
    const long long fz = blockIdx.z;

    static_assert(TILES == 1 || SBITS == 8, "several tiles per thread: 8-bit sources");
    RawTile8 raw[TILES];
    bool full[TILES];
#pragma unroll
    for (int t = 0; t < TILES; t++) {
        const int y0 = yb + 2 * t;
        full[t] = vec_ok && (x0 + 8 <= src.w) && (y0 + 2 <= src.h);
        if (TILES > 1 && full[t]) raw[t] = load_raw_tile8<L>(src, fz, x0, y0);
    }

    constexpr float YB = -(GMATB_MAGIC + (SBITS == 8 ? 16.f : 4096.f));
    constexpr float CB = -(GMATB_MAGIC + (SBITS == 8 ? 128.f : 32768.f));
    constexpr int BPP = dst_bpp(DST);
#pragma unroll
    for (int t = 0; t < TILES; t++) {
        const int y0 = yb + 2 * t;
        if (y0 >= src.h) break;
        float ym[2][8], um[4], vm[4];
        if (TILES > 1 && full[t]) unpack_raw_tile8<L>(raw[t], ym, um, vm);
        else load_yuv_tile<L, SBITS>(src, fz, x0, y0, full[t], ym, um, vm);
        int r[2][8], g[2][8], b[2][8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float fu, fv;
            upk(add2(pk(um[j], vm[j]), bc(CB)), fu, fv);
            ChromaTerms ct = chroma_terms<SPARSE, FMAF>(fu, fv, M);
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                f2 fy2 = add2(pk(ym[rr][2 * j], ym[rr][2 * j + 1]), bc(YB));
                csc_pair_i<SPARSE, FMAF>(fy2, ct, M, r[rr][2 * j], r[rr][2 * j + 1], g[rr][2 * j], g[rr][2 * j + 1],
                                         b[rr][2 * j], b[rr][2 * j + 1]);
            }
        }
        uint8_t *pd = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)y0 * dst.pl[0].pitch + (size_t)x0 * BPP;
        if (full[t]) {
            store_rgb_row8<DST, SBITS>(pd, r[0], g[0], b[0]);
            store_rgb_row8<DST, SBITS>(pd + dst.pl[0].pitch, r[1], g[1], b[1]);
        } else {
#pragma unroll
            for (int rr = 0; rr < 2; rr++)
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (x0 + i < src.w && y0 + rr < src.h)
                        store_rgb_px<DST, SBITS>(pd + (size_t)rr * dst.pl[0].pitch + i * BPP, r[rr][i], g[rr][i], b[rr][i]);
        }
    }
}

// yuv -> planar float rgb with (c - shift)/norm   (yuv2rgb_cuda.cu:381-433; format_cuda_kernel.cu:257-297)
// c is an integer in 0..255, so the IEEE division has only 3 x 256 possible results: each block computes them
// once into shared memory (768 __fdiv_rn per 8192 pixels instead of 3 per pixel -- the division expands to
// ~10 instructions and made this kernel compute-bound at 32 % of the HBM roofline) and every pixel is a
// byte-indexed lookup of the exact quotient.
template <int L, bool SPARSE>
__global__ void __launch_bounds__(256) yuv2rgb_planar_f32_kernel(Img src, Img dst, Mat9 M, float norm,
                                                                  float sr, float sg, float sb, int vec_ok) {
    __shared__ float tab[3][256];
    {
        const int t = threadIdx.y * 32 + threadIdx.x;
        tab[0][t] = __fdiv_rn(__fsub_rn((float)t, sr), norm);
        tab[1][t] = __fdiv_rn(__fsub_rn((float)t, sg), norm);
        tab[2][t] = __fdiv_rn(__fsub_rn((float)t, sb), norm);
    }
    __syncthreads();
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int ybase = (blockIdx.y * 8 + threadIdx.y) * 4;          // two row pairs per thread
    const long long fz = blockIdx.z;
    if (x0 >= src.w) return;
#pragma unroll 1
    for (int rp = 0; rp < 2; rp++) {
        const int y0 = ybase + 2 * rp;
        if (y0 >= src.h) return;
        const bool full = vec_ok && (x0 + 8 <= src.w) && (y0 + 2 <= src.h);
        float ym[2][8], um[4], vm[4];
        load_yuv_tile<L, 8>(src, fz, x0, y0, full, ym, um, vm);
        constexpr float YB = -(GMATB_MAGIC + 16.f), CB = -(GMATB_MAGIC + 128.f);
        int r[2][8], g[2][8], b[2][8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            float fu, fv;
            upk(add2(pk(um[j], vm[j]), bc(CB)), fu, fv);
            ChromaTerms t = chroma_terms<SPARSE, true>(fu, fv, M);     // planar kernels of the reference: FMA form
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                f2 fy2 = add2(pk(ym[rr][2 * j], ym[rr][2 * j + 1]), bc(YB));
                csc_pair_i<SPARSE, true>(fy2, t, M, r[rr][2 * j], r[rr][2 * j + 1], g[rr][2 * j], g[rr][2 * j + 1],
                                         b[rr][2 * j], b[rr][2 * j + 1]);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
            uint8_t *pd = dst.pl[c].p + fz * dst.pl[c].bstride + (size_t)y0 * dst.pl[c].pitch + (size_t)x0 * 4;
#pragma unroll
            for (int rr = 0; rr < 2; rr++) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; i++)
                    o[i] = tab[c][clamp_i(c == 0 ? r[rr][i] : c == 1 ? g[rr][i] : b[rr][i], 255)];
                float *q = reinterpret_cast<float *>(pd + (size_t)rr * dst.pl[c].pitch);
                if (full) {
                    __stcs(reinterpret_cast<float4 *>(q), make_float4(o[0], o[1], o[2], o[3]));
                    __stcs(reinterpret_cast<float4 *>(q) + 1, make_float4(o[4], o[5], o[6], o[7]));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++)
                        if (x0 + i < src.w && y0 + rr < src.h) q[i] = o[i];
                }
            }
        }
    }
}


// NV12 / P016 -> three stacked planes of 8-bit or float components, FMA form: metrans NvCodec/ColorSpace.cu
// YuvToRgbPlanarKernel (:163-190) -- the chain, clamp to the source range, 16-bit results >> 8, then the byte itself
// or v / 255 (IEEE division: a 256-entry table of the exact quotients per block).  SWAP: plane order B, G, R.
template <int SBITS, bool OUTF, bool SWAP, bool SPARSE>
__global__ void __launch_bounds__(256) yuv2rgb_planar8_kernel(Img src, uint8_t *dst, int dpitch, long long plane_stride, Mat9 M, int vec_ok) {
    __shared__ float tab[256];
    if (OUTF) {
        const int t = threadIdx.y * 32 + threadIdx.x;
        tab[t] = __fdiv_rn((float)t, 255.0f);
        __syncthreads();
    }
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (x0 >= src.w || y0 >= src.h) return;
    const bool full = vec_ok && (x0 + 8 <= src.w) && (y0 + 2 <= src.h);
    float ym[2][8], um[4], vm[4];
    load_yuv_tile<L_NV12, SBITS>(src, 0, x0, y0, full, ym, um, vm);
    constexpr float YB = -(GMATB_MAGIC + (SBITS == 8 ? 16.f : 4096.f));
    constexpr float CB = -(GMATB_MAGIC + (SBITS == 8 ? 128.f : 32768.f));
    constexpr int SMAX = SBITS == 8 ? 255 : 65535;
    int c[3][2][8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float fu, fv;
        upk(add2(pk(um[j], vm[j]), bc(CB)), fu, fv);
        ChromaTerms t = chroma_terms<SPARSE, true>(fu, fv, M);
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            f2 fy2 = add2(pk(ym[rr][2 * j], ym[rr][2 * j + 1]), bc(YB));
            csc_pair_i<SPARSE, true>(fy2, t, M, c[0][rr][2 * j], c[0][rr][2 * j + 1], c[1][rr][2 * j], c[1][rr][2 * j + 1],
                                     c[2][rr][2 * j], c[2][rr][2 * j + 1]);
        }
    }
#pragma unroll
    for (int p = 0; p < 3; p++) {
        const int ch = SWAP ? 2 - p : p;
        uint8_t *pd = dst + (size_t)p * plane_stride + (size_t)y0 * dpitch + (size_t)x0 * (OUTF ? 4 : 1);
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            int v[8];
#pragma unroll
            for (int i = 0; i < 8; i++) { v[i] = clamp_i(c[ch][rr][i], SMAX); if (SBITS == 16) v[i] >>= 8; }
            uint8_t *q = pd + (size_t)rr * dpitch;
            if (OUTF) {
                float *qf = reinterpret_cast<float *>(q);
                if (full && ((((uintptr_t)dst | (uintptr_t)dpitch | (uintptr_t)plane_stride) & 15) == 0)) {
                    __stcs(reinterpret_cast<float4 *>(qf), make_float4(tab[v[0]], tab[v[1]], tab[v[2]], tab[v[3]]));
                    __stcs(reinterpret_cast<float4 *>(qf) + 1, make_float4(tab[v[4]], tab[v[5]], tab[v[6]], tab[v[7]]));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) if (x0 + i < src.w && y0 + rr < src.h) qf[i] = tab[v[i]];
                }
            } else {
                if (full && ((((uintptr_t)dst | (uintptr_t)dpitch | (uintptr_t)plane_stride) & 7) == 0)) {
                    stg64(q, make_uint2(pack4_u8(v[0], v[1], v[2], v[3]), pack4_u8(v[4], v[5], v[6], v[7])));
                } else {
#pragma unroll
                    for (int i = 0; i < 8; i++) if (x0 + i < src.w && y0 + rr < src.h) q[i] = (uint8_t)v[i];
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// 8-sample / 4-sample vector row accesses (samples of 1 or 2 bytes)
// ---------------------------------------------------------------------------
template <int BYTES> __device__ __forceinline__ void load8(const uint8_t *p, unsigned (&o)[8]) {
    if (BYTES == 1) {
        uint2 w = ldg64(p);
#pragma unroll
        for (int i = 0; i < 4; i++) { o[i] = (w.x >> (8 * i)) & 0xFFu; o[4 + i] = (w.y >> (8 * i)) & 0xFFu; }
    } else {
        uint4 w = ldg128(p);
        o[0] = w.x & 0xFFFFu; o[1] = w.x >> 16; o[2] = w.y & 0xFFFFu; o[3] = w.y >> 16;
        o[4] = w.z & 0xFFFFu; o[5] = w.z >> 16; o[6] = w.w & 0xFFFFu; o[7] = w.w >> 16;
    }
}
template <int BYTES> __device__ __forceinline__ void store8(uint8_t *p, const unsigned (&v)[8]) {
    if (BYTES == 1) {
        stg64(p, make_uint2(v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24), v[4] | (v[5] << 8) | (v[6] << 16) | (v[7] << 24)));
    } else {
        stg128(p, make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16)));
    }
}
template <int BYTES> __device__ __forceinline__ void load4(const uint8_t *p, unsigned (&o)[4]) {
    if (BYTES == 1) {
        uint32_t w = ldg32(p);
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (w >> (8 * i)) & 0xFFu;
    } else {
        uint2 w = ldg64(p);
        o[0] = w.x & 0xFFFFu; o[1] = w.x >> 16; o[2] = w.y & 0xFFFFu; o[3] = w.y >> 16;
    }
}
template <int BYTES> __device__ __forceinline__ void store4(uint8_t *p, const unsigned (&v)[4]) {
    if (BYTES == 1) stg32(p, v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24));
    else stg64(p, make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16)));
}
// chroma of a 4:2:0 tile (4 samples each of U and V), either layout
template <int L, int BYTES>
__device__ __forceinline__ void load_chroma4(const Img &s, long long fz, int cx, int cy, unsigned (&u)[4], unsigned (&v)[4]) {
    if (L == L_NV12) {
        unsigned t[8];
        load8<BYTES>(s.pl[1].p + fz * s.pl[1].bstride + (size_t)cy * s.pl[1].pitch + (size_t)cx * 2 * BYTES, t);
#pragma unroll
        for (int j = 0; j < 4; j++) { u[j] = t[2 * j]; v[j] = t[2 * j + 1]; }
    } else {
        load4<BYTES>(s.pl[1].p + fz * s.pl[1].bstride + (size_t)cy * s.pl[1].pitch + (size_t)cx * BYTES, u);
        load4<BYTES>(s.pl[2].p + fz * s.pl[2].bstride + (size_t)cy * s.pl[2].pitch + (size_t)cx * BYTES, v);
    }
}
template <int L, int BYTES>
__device__ __forceinline__ void store_chroma4(const Img &d, long long fz, int cx, int cy, const unsigned (&u)[4], const unsigned (&v)[4]) {
    if (L == L_NV12) {
        unsigned t[8];
#pragma unroll
        for (int j = 0; j < 4; j++) { t[2 * j] = u[j]; t[2 * j + 1] = v[j]; }
        store8<BYTES>(d.pl[1].p + fz * d.pl[1].bstride + (size_t)cy * d.pl[1].pitch + (size_t)cx * 2 * BYTES, t);
    } else {
        store4<BYTES>(d.pl[1].p + fz * d.pl[1].bstride + (size_t)cy * d.pl[1].pitch + (size_t)cx * BYTES, u);
        store4<BYTES>(d.pl[2].p + fz * d.pl[2].bstride + (size_t)cy * d.pl[2].pitch + (size_t)cx * BYTES, v);
    }
}

// ---------------------------------------------------------------------------
// packed rgb -> yuv 4:2:0   (yuv2rgb_cuda.cu:653-739)
// ---------------------------------------------------------------------------
enum { S_RGB24 = 0, S_BGR24, S_RGBA, S_BGRA, S_RGBA64, S_BGRA64 };
__host__ __device__ constexpr int srgb_bpp(int s) { return s <= S_BGR24 ? 3 : s <= S_BGRA ? 4 : 8; }
__host__ __device__ constexpr bool srgb_swap(int s) { return s == S_BGR24 || s == S_BGRA || s == S_BGRA64; }
__host__ __device__ constexpr bool srgb_is16(int s) { return s >= S_RGBA64; }

// 8-bit packed rgb -> 8-bit yuv 4:2:0, one full 8x2 tile: the vector path of rgb2yuv_kernel.
// Same IEEE operations as rgb2y_f / the reference (FADD(FFMA(b,m2,FFMA(r,m0,FMUL(g,m1))),low), integer 2x2
// mean for chroma), but nothing touches the 16-lane conversion unit: bytes become floats through PRMT magic
// numbers (2^23 + byte, minus 2^23 is exact), the luma chain runs packed on (top,bottom) pairs, U and V of a
// block run as one packed pair, the 2x2 mean is exact float arithmetic (sum <= 1020, x 0.25, floor through
// RZ(q + 2^23)), truncation is the round-toward-zero multiply by 2^-149 and the clamp is the saturating I2IP
// pack.  (The I2F form ran at 55 % of the HBM roofline, conversion-unit bound.)
template <int SRC, int L>
__device__ __forceinline__ void rgb2yuv_tile8_load(const Img &src, long long fz, int x0, int y0, uint32_t (&w)[2][2 * srgb_bpp(SRC)]) {
    constexpr int BPP = srgb_bpp(SRC);
    constexpr int NW = 2 * BPP;                       // 32-bit words per 8-pixel row
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        const uint8_t *row = src.pl[0].p + fz * src.pl[0].bstride + (size_t)(y0 + rr) * src.pl[0].pitch + (size_t)x0 * BPP;
        if (BPP == 4) {
            const uint4 a = ldg128(row), b = ldg128(row + 16);
            w[rr][0] = a.x; w[rr][1] = a.y; w[rr][2] = a.z; w[rr][3] = a.w;
            w[rr][NW - 4] = b.x; w[rr][NW - 3] = b.y; w[rr][NW - 2] = b.z; w[rr][NW - 1] = b.w;
        } else {
            const uint2 a = ldg64(row), b = ldg64(row + 8), c = ldg64(row + 16);
            w[rr][0] = a.x; w[rr][1] = a.y; w[rr][2] = b.x; w[rr][3] = b.y; w[rr][4] = c.x; w[rr][5] = c.y;
        }
    }
}
template <int SRC, int L>
__device__ __forceinline__ void rgb2yuv_tile8_convert(const uint32_t (&w)[2][2 * srgb_bpp(SRC)], const Img &dst, const Mat9 &M,
                                                      long long fz, int x0, int y0) {
    constexpr int BPP = srgb_bpp(SRC);
    const f2 nm = bc(-GMATB_MAGIC), z = bc(GMATB_TWO_M149);
    f2 c2[3][8];                                      // [component r,g,b][column] = (top, bottom), exact integers as floats
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int bi = i * BPP + k, comp = srgb_swap(SRC) ? 2 - k : k;
            const float t = __uint_as_float(__byte_perm(w[0][bi >> 2], 0x4B000000u, 0x7440u | (bi & 3)));
            const float b = __uint_as_float(__byte_perm(w[1][bi >> 2], 0x4B000000u, 0x7440u | (bi & 3)));
            c2[comp][i] = add2(pk(t, b), nm);
        }
    int yt[8], yb[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        f2 t = mul2(c2[1][i], bc(M.m[1]));
        t = fma2(c2[0][i], bc(M.m[0]), t);
        t = fma2(c2[2][i], bc(M.m[2]), t);
        upki(mul2_rz(add2(t, bc(16.0f)), z), yt[i], yb[i]);
    }
    uint8_t *pdy = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)y0 * dst.pl[0].pitch + x0;
    stg64(pdy, make_uint2(pack4_u8(yt[0], yt[1], yt[2], yt[3]), pack4_u8(yt[4], yt[5], yt[6], yt[7])));
    stg64(pdy + dst.pl[0].pitch, make_uint2(pack4_u8(yb[0], yb[1], yb[2], yb[3]), pack4_u8(yb[4], yb[5], yb[6], yb[7])));
    int uu[4], vv[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        float mean[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            float st, sb;
            upk(add2(c2[k][2 * j], c2[k][2 * j + 1]), st, sb);             // row sums, then the block sum: all exact
            const float q = __fmul_rn(__fadd_rn(st, sb), 0.25f);
            mean[k] = __fadd_rn(__fadd_rz(q, GMATB_MAGIC), -GMATB_MAGIC);    // (r0+r1+r2+r3)/4 in integers (:685-687)
        }
        f2 t = mul2(bc(mean[1]), pk(M.m[4], M.m[7]));
        t = fma2(bc(mean[0]), pk(M.m[3], M.m[6]), t);
        t = fma2(bc(mean[2]), pk(M.m[5], M.m[8]), t);
        upki(mul2_rz(add2(t, bc(128.0f)), z), uu[j], vv[j]);
    }
    const int cx = x0 >> 1, cy = y0 >> 1;
    if (L == L_NV12) {
        stg64(dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + (size_t)cx * 2,
              make_uint2(pack4_u8(uu[0], vv[0], uu[1], vv[1]), pack4_u8(uu[2], vv[2], uu[3], vv[3])));
    } else {
        stg32(dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + cx, pack4_u8(uu[0], uu[1], uu[2], uu[3]));
        stg32(dst.pl[2].p + fz * dst.pl[2].bstride + (size_t)cy * dst.pl[2].pitch + cx, pack4_u8(vv[0], vv[1], vv[2], vv[3]));
    }
}

// DBITS: 8 (NV12 / YUV420P) or 16 (P016 from 64-bit rgb, yuv2rgb_cuda.cu:741-746).
// 16-bit rgb -> 8-bit yuv uses the high byte of each component (the reference
// mis-reads every non-RGB24 source as RGB24, :748-762 -- not reproduced).
template <int SRC, int L, int DBITS>
#ifndef GMATB_R2Y_MINB
#define GMATB_R2Y_MINB 3
#endif
__device__ __forceinline__ void rgb2yuv_tile(const Img &src, const Img &dst, const Mat9 &M, int vec_ok, long long fz, int x0, int y0) {
    const int W = src.w, H = src.h;
    constexpr int BPP = srgb_bpp(SRC);
    constexpr float LOW = DBITS == 8 ? 16.f : 4096.f, MID = DBITS == 8 ? 128.f : 32768.f;
    constexpr int DMAX = DBITS == 8 ? 255 : 65535;
    constexpr int DBS = DBITS / 8;
    const uint8_t *ps = src.pl[0].p + fz * src.pl[0].bstride;

    const bool full = vec_ok && x0 + 8 <= W && y0 + 2 <= H;
    if constexpr (!srgb_is16(SRC) && DBITS == 8) {
        if (full) {
            uint32_t w[2][2 * BPP];
            rgb2yuv_tile8_load<SRC, L>(src, fz, x0, y0, w);
            rgb2yuv_tile8_convert<SRC, L>(w, dst, M, fz, x0, y0);
            return;
        }
    }
    int cr[2][8], cg[2][8], cb[2][8];
    if (full) {
        // 128/64-bit row loads: 24 / 32 / 64 bytes per 8-pixel row
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            const uint8_t *row = ps + (size_t)(y0 + rr) * src.pl[0].pitch + (size_t)x0 * BPP;
            unsigned a[8], b[8], c[8];
            if (srgb_is16(SRC)) {
#pragma unroll
                for (int h = 0; h < 4; h++) {
                    uint4 w = ldg128(row + 16 * h);
                    a[2 * h] = w.x & 0xFFFFu; b[2 * h] = w.x >> 16; c[2 * h] = w.y & 0xFFFFu;
                    a[2 * h + 1] = w.z & 0xFFFFu; b[2 * h + 1] = w.z >> 16; c[2 * h + 1] = w.w & 0xFFFFu;
                }
                if (DBITS == 8) {
#pragma unroll
                    for (int i = 0; i < 8; i++) { a[i] >>= 8; b[i] >>= 8; c[i] >>= 8; }
                }
            } else if (BPP == 4) {
                uint4 w0 = ldg128(row), w1 = ldg128(row + 16);
                const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) { a[i] = w[i] & 0xFFu; b[i] = (w[i] >> 8) & 0xFFu; c[i] = (w[i] >> 16) & 0xFFu; }
            } else {
                uint2 q0 = ldg64(row), q1 = ldg64(row + 8), q2 = ldg64(row + 16);
                const uint32_t w[6] = {q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    unsigned t[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) { const int bi = 3 * i + k; t[k] = (w[bi >> 2] >> (8 * (bi & 3))) & 0xFFu; }
                    a[i] = t[0]; b[i] = t[1]; c[i] = t[2];
                }
            }
#pragma unroll
            for (int i = 0; i < 8; i++) { cr[rr][i] = srgb_swap(SRC) ? c[i] : a[i]; cg[rr][i] = b[i]; cb[rr][i] = srgb_swap(SRC) ? a[i] : c[i]; }
        }
    } else
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        int yy = min(y0 + rr, H - 1);
        const uint8_t *row = ps + (size_t)yy * src.pl[0].pitch;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            int xx = min(x0 + i, W - 1);
            int a, b, c;
            if (srgb_is16(SRC)) {
                ushort4 v = __ldcs(reinterpret_cast<const ushort4 *>(row + (size_t)xx * 8));
                a = v.x; b = v.y; c = v.z;
                if (DBITS == 8) { a >>= 8; b >>= 8; c >>= 8; }
            } else if (BPP == 4) {
                uchar4 v = __ldcs(reinterpret_cast<const uchar4 *>(row + (size_t)xx * 4));
                a = v.x; b = v.y; c = v.z;
            } else {
                const uint8_t *q = row + (size_t)xx * 3;
                a = q[0]; b = q[1]; c = q[2];
            }
            cr[rr][i] = srgb_swap(SRC) ? c : a; cg[rr][i] = b; cb[rr][i] = srgb_swap(SRC) ? a : c;
        }
    }
    uint8_t *pdy = dst.pl[0].p + fz * dst.pl[0].bstride;
    if (full) {
        unsigned uu[4], vv[4];
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            unsigned yv[8];
#pragma unroll
            for (int i = 0; i < 8; i++)
                yv[i] = clamp_i(trunc_i(rgb2y_f((float)cr[rr][i], (float)cg[rr][i], (float)cb[rr][i], M, 0, LOW)), DMAX);
            store8<DBS>(pdy + (size_t)(y0 + rr) * dst.pl[0].pitch + (size_t)x0 * DBS, yv);
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int mr = (cr[0][2 * j] + cr[0][2 * j + 1] + cr[1][2 * j] + cr[1][2 * j + 1]) / 4;
            const int mg = (cg[0][2 * j] + cg[0][2 * j + 1] + cg[1][2 * j] + cg[1][2 * j + 1]) / 4;
            const int mb = (cb[0][2 * j] + cb[0][2 * j + 1] + cb[1][2 * j] + cb[1][2 * j + 1]) / 4;
            uu[j] = clamp_i(trunc_i(rgb2y_f((float)mr, (float)mg, (float)mb, M, 1, MID)), DMAX);
            vv[j] = clamp_i(trunc_i(rgb2y_f((float)mr, (float)mg, (float)mb, M, 2, MID)), DMAX);
        }
        store_chroma4<L, DBS>(dst, fz, x0 >> 1, y0 >> 1, uu, vv);
        return;
    }
#pragma unroll
    for (int rr = 0; rr < 2; rr++)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (x0 + i < W && y0 + rr < H) {
                int v = clamp_i(trunc_i(rgb2y_f((float)cr[rr][i], (float)cg[rr][i], (float)cb[rr][i], M, 0, LOW)), DMAX);
                uint8_t *q = pdy + (size_t)(y0 + rr) * dst.pl[0].pitch + (size_t)(x0 + i) * DBS;
                if (DBITS == 8) *q = v; else *reinterpret_cast<uint16_t *>(q) = v;
            }
        }
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (x0 + 2 * j >= W) break;
        // integer mean of the 2x2 block in the component type (:685-687)
        int mr = (cr[0][2 * j] + cr[0][2 * j + 1] + cr[1][2 * j] + cr[1][2 * j + 1]) / 4;
        int mg = (cg[0][2 * j] + cg[0][2 * j + 1] + cg[1][2 * j] + cg[1][2 * j + 1]) / 4;
        int mb = (cb[0][2 * j] + cb[0][2 * j + 1] + cb[1][2 * j] + cb[1][2 * j + 1]) / 4;
        int u = clamp_i(trunc_i(rgb2y_f((float)mr, (float)mg, (float)mb, M, 1, MID)), DMAX);
        int v = clamp_i(trunc_i(rgb2y_f((float)mr, (float)mg, (float)mb, M, 2, MID)), DMAX);
        int cx = (x0 >> 1) + j, cy = y0 >> 1;
        if (L == L_NV12) {
            uint8_t *q = dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + (size_t)cx * 2 * DBS;
            if (DBITS == 8) { q[0] = u; q[1] = v; }
            else { reinterpret_cast<uint16_t *>(q)[0] = u; reinterpret_cast<uint16_t *>(q)[1] = v; }
        } else {
            uint8_t *qu = dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + (size_t)cx * DBS;
            uint8_t *qv = dst.pl[2].p + fz * dst.pl[2].bstride + (size_t)cy * dst.pl[2].pitch + (size_t)cx * DBS;
            if (DBITS == 8) { *qu = u; *qv = v; }
            else { *reinterpret_cast<uint16_t *>(qu) = u; *reinterpret_cast<uint16_t *>(qv) = v; }
        }
    }
}

// 8-bit rgb -> 8-bit yuv runs two row pairs per thread with the loads of both issued first (96 bytes in flight per thread: the
// one-pair form measured 74 % of the HBM roofline, latency-bound); every other instantiation one pair.
template <int SRC, int DBITS> __host__ __device__ constexpr int r2y_tiles() { return (!srgb_is16(SRC) && DBITS == 8) ? 2 : 1; }
template <int SRC, int L, int DBITS>
__global__ void __launch_bounds__(256, (r2y_tiles<SRC, DBITS>() > 1 ? 2 : GMATB_R2Y_MINB)) rgb2yuv_kernel(Img src, Img dst, Mat9 M, int vec_ok) {
    constexpr int T = r2y_tiles<SRC, DBITS>();
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int yb = (blockIdx.y * 8 + threadIdx.y) * 2 * T;
    if (x0 >= src.w || yb >= src.h) return;
    const long long fz = blockIdx.z;
    if constexpr (T == 1) {
        rgb2yuv_tile<SRC, L, DBITS>(src, dst, M, vec_ok, fz, x0, yb);
    } else {
        uint32_t w[T][2][2 * srgb_bpp(SRC)];
        bool full[T];
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int y0 = yb + 2 * t;
            full[t] = vec_ok && x0 + 8 <= src.w && y0 + 2 <= src.h;
            if (full[t]) rgb2yuv_tile8_load<SRC, L>(src, fz, x0, y0, w[t]);
        }
#pragma unroll
        for (int t = 0; t < T; t++) {
            const int y0 = yb + 2 * t;
            if (y0 >= src.h) break;
            if (full[t]) rgb2yuv_tile8_convert<SRC, L>(w[t], dst, M, fz, x0, y0);
            else rgb2yuv_tile<SRC, L, DBITS>(src, dst, M, 0, fz, x0, y0);
        }
    }
}

// ---------------------------------------------------------------------------
// yuv 4:2:0 repack / bit-depth change   (yuv2yuv_cuda.cu:56-63, :194-286)
// depth codes: 8, 10 (MSB-aligned in 16 bits, as P010), 16
// ---------------------------------------------------------------------------
template <int SD, int DD> __device__ __forceinline__ unsigned conv_depth(unsigned x) {
    if (SD == DD) return x;
    if (SD == 8 && DD == 10) return (x | (x << 8)) & 0xFFC0u;
    if (SD == 8 && DD == 16) return (x | (x << 8)) & 0xFFFFu;
    if (DD == 8) return x >> 8;
    if (SD == 10 && DD == 16) return x | (x >> 10);
    return x & 0xFFC0u;   // 16 -> 10
}

// Same depth on both sides (nv12 <-> yuv420p, p016 <-> yuv420p16 ...): nothing is converted, so the luma plane is a
// straight 128-bit copy and the chroma rows are (de)interleaved with byte permutes.  A thread moves 16 luma bytes of two
// rows and the 16 chroma bytes under them.  (The element-wise kernel below ran this case at 76 % of the measured copy peak.)
template <int SL, int DL, int SB>
__global__ void __launch_bounds__(256) yuv2yuv_copy_kernel(Img src, Img dst) {
    const int xb = (blockIdx.x * 32 + threadIdx.x) * 16;                 // byte offset in a luma row
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    if (xb >= src.w * SB || y0 >= src.h) return;
    const long long fz = blockIdx.z;
    const uint8_t *py = src.pl[0].p + fz * src.pl[0].bstride + (size_t)y0 * src.pl[0].pitch + xb;
    uint8_t *qy = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)y0 * dst.pl[0].pitch + xb;
    const uint4 a = ldg128(py), b = ldg128(py + src.pl[0].pitch);
    const int cy = y0 >> 1;
    uint2 u, v;                                                         // 8 bytes of U, 8 bytes of V
    if (SL == L_NV12) {
        const uint4 c = ldg128(src.pl[1].p + fz * src.pl[1].bstride + (size_t)cy * src.pl[1].pitch + xb);
        if (SB == 1) {
            u = make_uint2(prmt(c.x, c.y, 0x6420u), prmt(c.z, c.w, 0x6420u));
            v = make_uint2(prmt(c.x, c.y, 0x7531u), prmt(c.z, c.w, 0x7531u));
        } else {
            u = make_uint2(prmt(c.x, c.y, 0x5410u), prmt(c.z, c.w, 0x5410u));
            v = make_uint2(prmt(c.x, c.y, 0x7632u), prmt(c.z, c.w, 0x7632u));
        }
    } else {
        u = ldg64(src.pl[1].p + fz * src.pl[1].bstride + (size_t)cy * src.pl[1].pitch + (xb >> 1));
        v = ldg64(src.pl[2].p + fz * src.pl[2].bstride + (size_t)cy * src.pl[2].pitch + (xb >> 1));
    }
    stg128(qy, a); stg128(qy + dst.pl[0].pitch, b);
    if (DL == L_NV12) {
        uint4 c;
        if (SB == 1) { c.x = prmt(u.x, v.x, 0x5140u); c.y = prmt(u.x, v.x, 0x7362u); c.z = prmt(u.y, v.y, 0x5140u); c.w = prmt(u.y, v.y, 0x7362u); }
        else         { c.x = prmt(u.x, v.x, 0x5410u); c.y = prmt(u.x, v.x, 0x7632u); c.z = prmt(u.y, v.y, 0x5410u); c.w = prmt(u.y, v.y, 0x7632u); }
        stg128(dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + xb, c);
    } else {
        stg64(dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + (xb >> 1), u);
        stg64(dst.pl[2].p + fz * dst.pl[2].bstride + (size_t)cy * dst.pl[2].pitch + (xb >> 1), v);
    }
}

template <int SL, int SD, int DL, int DD>
__global__ void __launch_bounds__(256) yuv2yuv_kernel(Img src, Img dst, int vec_ok) {
    constexpr int SB = SD == 8 ? 1 : 2, DB = DD == 8 ? 1 : 2;
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int y0 = (blockIdx.y * 8 + threadIdx.y) * 2;
    const int W = src.w, H = src.h;
    if (x0 >= W || y0 >= H) return;
    const long long fz = blockIdx.z;
    const int cw = (W + 1) >> 1;
    if (vec_ok && x0 + 8 <= W && y0 + 2 <= H) {
#pragma unroll
        for (int rr = 0; rr < 2; rr++) {
            unsigned yv[8];
            load8<SB>(src.pl[0].p + fz * src.pl[0].bstride + (size_t)(y0 + rr) * src.pl[0].pitch + (size_t)x0 * SB, yv);
#pragma unroll
            for (int i = 0; i < 8; i++) yv[i] = conv_depth<SD, DD>(yv[i]);
            store8<DB>(dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)(y0 + rr) * dst.pl[0].pitch + (size_t)x0 * DB, yv);
        }
        unsigned u[4], v[4];
        load_chroma4<SL, SB>(src, fz, x0 >> 1, y0 >> 1, u, v);
#pragma unroll
        for (int j = 0; j < 4; j++) { u[j] = conv_depth<SD, DD>(u[j]); v[j] = conv_depth<SD, DD>(v[j]); }
        store_chroma4<DL, DB>(dst, fz, x0 >> 1, y0 >> 1, u, v);
        return;
    }
#pragma unroll
    for (int rr = 0; rr < 2; rr++) {
        if (y0 + rr >= H) break;
        const uint8_t *ps = src.pl[0].p + fz * src.pl[0].bstride + (size_t)(y0 + rr) * src.pl[0].pitch + (size_t)x0 * SB;
        uint8_t *pd = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)(y0 + rr) * dst.pl[0].pitch + (size_t)x0 * DB;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (x0 + i >= W) break;
            unsigned v = SB == 1 ? ps[i] : reinterpret_cast<const uint16_t *>(ps)[i];
            v = conv_depth<SD, DD>(v);
            if (DB == 1) pd[i] = v; else reinterpret_cast<uint16_t *>(pd)[i] = v;
        }
    }
    const int cy = y0 >> 1;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int cx = (x0 >> 1) + j;
        if (cx >= cw) break;
        unsigned u, v;
        if (SL == L_NV12) {
            const uint8_t *q = src.pl[1].p + fz * src.pl[1].bstride + (size_t)cy * src.pl[1].pitch + (size_t)cx * 2 * SB;
            if (SB == 1) { u = q[0]; v = q[1]; } else { u = reinterpret_cast<const uint16_t *>(q)[0]; v = reinterpret_cast<const uint16_t *>(q)[1]; }
        } else {
            const uint8_t *qu = src.pl[1].p + fz * src.pl[1].bstride + (size_t)cy * src.pl[1].pitch + (size_t)cx * SB;
            const uint8_t *qv = src.pl[2].p + fz * src.pl[2].bstride + (size_t)cy * src.pl[2].pitch + (size_t)cx * SB;
            if (SB == 1) { u = *qu; v = *qv; } else { u = *reinterpret_cast<const uint16_t *>(qu); v = *reinterpret_cast<const uint16_t *>(qv); }
        }
        u = conv_depth<SD, DD>(u); v = conv_depth<SD, DD>(v);
        if (DL == L_NV12) {
            uint8_t *q = dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + (size_t)cx * 2 * DB;
            if (DB == 1) { q[0] = u; q[1] = v; } else { reinterpret_cast<uint16_t *>(q)[0] = u; reinterpret_cast<uint16_t *>(q)[1] = v; }
        } else {
            uint8_t *qu = dst.pl[1].p + fz * dst.pl[1].bstride + (size_t)cy * dst.pl[1].pitch + (size_t)cx * DB;
            uint8_t *qv = dst.pl[2].p + fz * dst.pl[2].bstride + (size_t)cy * dst.pl[2].pitch + (size_t)cx * DB;
            if (DB == 1) { *qu = u; *qv = v; } else { *reinterpret_cast<uint16_t *>(qu) = u; *reinterpret_cast<uint16_t *>(qv) = v; }
        }
    }
}

// ---------------------------------------------------------------------------
// rgb24 <-> bgr24   (rgb2rgb_cuda_kernel.cu:6-22): 8 pixels (24 bytes) per thread
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rgb24_swap_kernel(Img src, Img dst, int vec_ok) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 8;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x0 >= src.w || y >= src.h) return;
    const long long fz = blockIdx.z;
    const uint8_t *ps = src.pl[0].p + fz * src.pl[0].bstride + (size_t)y * src.pl[0].pitch + (size_t)x0 * 3;
    uint8_t *pd = dst.pl[0].p + fz * dst.pl[0].bstride + (size_t)y * dst.pl[0].pitch + (size_t)x0 * 3;
    if (vec_ok && x0 + 8 <= src.w) {
        uint2 a = ldg64(ps), b = ldg64(ps + 8), c = ldg64(ps + 16);
        uint32_t w[6] = {a.x, a.y, b.x, b.y, c.x, c.y}, o[6];
        // bytes: R0G0B0R1 G1B1R2G2 B2R3G3B3 | same pattern for pixels 4..7
#pragma unroll
        for (int h = 0; h < 2; h++) {
            uint32_t w0 = w[3 * h], w1 = w[3 * h + 1], w2 = w[3 * h + 2];
            o[3 * h]     = prmt(w0, w1, 0x5012u);              // B0 G0 R0 B1   (B1 = w1.byte1 = idx5)
            // word1 = G1 R1 B2 G2 : G1=w1.b0, R1=w0.b3, B2=w2.b0, G2=w1.b3
            uint32_t t = prmt(w1, w0, 0x3070u);                // {w1.b0, w0.b3, w1.b0, w1.b3}
            o[3 * h + 1] = prmt(t, w2, 0x3410u);               // {t.b0, t.b1, w2.b0, t.b3}
            // word2 = R2 B3 G3 R3 : R2=w1.b2, B3=w2.b3, G3=w2.b2, R3=w2.b1
            o[3 * h + 2] = prmt(w2, w1, 0x1236u);              // {w1.b2, w2.b3, w2.b2, w2.b1}
        }
        stg64(pd, make_uint2(o[0], o[1])); stg64(pd + 8, make_uint2(o[2], o[3])); stg64(pd + 16, make_uint2(o[4], o[5]));
    } else {
        for (int i = 0; i < 8 && x0 + i < src.w; i++) {
            uint8_t r = ps[3 * i], g = ps[3 * i + 1], b = ps[3 * i + 2];
            pd[3 * i] = b; pd[3 * i + 1] = g; pd[3 * i + 2] = r;
        }
    }
}

// ---------------------------------------------------------------------------
// host dispatch
// ---------------------------------------------------------------------------
static inline dim3 tile_grid(int w, int h, int rows_per_thread, int batch) {
    return dim3((w + 255) / 256, (h + 8 * rows_per_thread - 1) / (8 * rows_per_thread), batch > 1 ? batch : 1);
}

static bool aligned16(const Img &a, int np) {
    for (int i = 0; i < np; i++)
        if (((uintptr_t)a.pl[i].p | (uintptr_t)a.pl[i].pitch | (uintptr_t)a.pl[i].bstride) & 15) return false;
    return true;
}

static int dst_code(int fmt) {
    switch (fmt) {
    case GMATB_FMT_RGB24: return D_RGB24;   case GMATB_FMT_BGR24: return D_BGR24;
    case GMATB_FMT_RGBA: return D_RGBA;     case GMATB_FMT_BGRA: return D_BGRA;
    case GMATB_FMT_RGB0: return D_RGBA;     case GMATB_FMT_BGR0: return D_BGRA;
    case GMATB_FMT_RGB48LE: return D_RGB48; case GMATB_FMT_BGR48LE: return D_BGR48;
    case GMATB_FMT_RGBA64LE: return D_RGBA64; case GMATB_FMT_BGRA64LE: return D_BGRA64;
    default: return -1;
    }
}

#define Y2R_TILES(SBITS) ((SBITS) == 8 ? 2 : 1)
template <int L, int SBITS, bool SPARSE>
static int launch_yuv2rgb_dst(int dc, dim3 g, cudaStream_t st, const Img &s, const Img &d, const Mat9 &M, int vec) {
    dim3 b(32, 8);
    switch (dc) {
#define C(D) case D: yuv2rgb_kernel<L, SBITS, D, SPARSE, (SBITS == 16), Y2R_TILES(SBITS)><<<g, b, 0, st>>>(s, d, M, vec); break;
        C(D_RGB24) C(D_BGR24) C(D_RGBA) C(D_BGRA) C(D_RGB48) C(D_BGR48) C(D_RGBA64) C(D_BGRA64)
#undef C
    default: return GMATB_ERR_UNSUPPORTED;
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

int yuv2rgb_planar8_launch(const GmatbImage *src, uint8_t *dst, int dpitch, long long plane_stride, bool outf, bool swap,
                           const Mat9 &M, cudaStream_t st) {
    if (!src || !dst || src->width <= 0 || src->height <= 0) return GMATB_ERR_INVAL;
    const bool b16 = src->format == GMATB_FMT_P016LE || src->format == GMATB_FMT_P010LE;
    if (!b16 && src->format != GMATB_FMT_NV12) return GMATB_ERR_UNSUPPORTED;
    Img s;
    if (!to_img(src, &s, 2)) return GMATB_ERR_INVAL;
    const bool sparse = (M.m[1] == 0.f && M.m[8] == 0.f);
    const int vec = aligned16(s, 2);
    dim3 g = tile_grid(s.w, s.h, 2, 1), b(32, 8);
#define K4(B, F, W) do { if (sparse) yuv2rgb_planar8_kernel<B, F, W, true><<<g, b, 0, st>>>(s, dst, dpitch, plane_stride, M, vec); \
                         else yuv2rgb_planar8_kernel<B, F, W, false><<<g, b, 0, st>>>(s, dst, dpitch, plane_stride, M, vec); } while (0)
#define K3(B, F) do { if (swap) K4(B, F, true); else K4(B, F, false); } while (0)
#define K2(B) do { if (outf) K3(B, true); else K3(B, false); } while (0)
    if (b16) K2(16); else K2(8);
#undef K2
#undef K3
#undef K4
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

// NV12 -> BGRA / RGBA / BGRA64 in the FMA form (metrans NvCodec/ColorSpace.cu: Nv12ToBgra32, Nv12ToRgba32, Nv12ToBgra64)
int yuv2rgb_nv12_fma_launch(const GmatbImage *src, const GmatbImage *dst, const Mat9 &M, cudaStream_t st) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height || src->width <= 0 || src->height <= 0 ||
        src->format != GMATB_FMT_NV12) return GMATB_ERR_INVAL;
    const int dc = dst_code(dst->format);
    Img s, d;
    if (!to_img(src, &s, 2) || !to_img(dst, &d, 1)) return GMATB_ERR_INVAL;
    const bool sparse = (M.m[1] == 0.f && M.m[8] == 0.f);
    const int vec = aligned16(s, 2) && aligned16(d, 1);
    dim3 g = tile_grid(s.w, s.h, 2, src->batch), b(32, 8);
#define GO(D) do { if (sparse) yuv2rgb_kernel<L_NV12, 8, D, true, true><<<g, b, 0, st>>>(s, d, M, vec); \
                   else yuv2rgb_kernel<L_NV12, 8, D, false, true><<<g, b, 0, st>>>(s, d, M, vec); } while (0)
    switch (dc) {
    case D_BGRA: GO(D_BGRA); break;
    case D_RGBA: GO(D_RGBA); break;
    case D_BGRA64: GO(D_BGRA64); break;
    default: return GMATB_ERR_UNSUPPORTED;
    }
#undef GO
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

int yuv2rgb_launch(const GmatbImage *src, const GmatbImage *dst, const Mat9 &M, cudaStream_t st) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height || src->width <= 0 || src->height <= 0)
        return GMATB_ERR_INVAL;
    int dc = dst_code(dst->format);
    if (dc < 0) return GMATB_ERR_UNSUPPORTED;
    Img s, d;
    if (!to_img(src, &s, fmt_planes(src->format)) || !to_img(dst, &d, 1)) return GMATB_ERR_INVAL;
    const bool sparse = (M.m[1] == 0.f && M.m[8] == 0.f);
    const int vec = aligned16(s, fmt_planes(src->format)) && aligned16(d, 1);
#define GO(L, B) (sparse ? launch_yuv2rgb_dst<L, B, true>(dc, tile_grid(s.w, s.h, 2 * Y2R_TILES(B), src->batch), st, s, d, M, vec) \
                        : launch_yuv2rgb_dst<L, B, false>(dc, tile_grid(s.w, s.h, 2 * Y2R_TILES(B), src->batch), st, s, d, M, vec))
    switch (src->format) {
    case GMATB_FMT_NV12:    return GO(L_NV12, 8);
    case GMATB_FMT_YUV420P: return GO(L_I420, 8);
    case GMATB_FMT_P010LE:
    case GMATB_FMT_P016LE:  return GO(L_NV12, 16);
    case GMATB_FMT_YUV420P10LE:   // container-aligned 16-bit planar, treated as the reference treats Unit10b/16b
    case GMATB_FMT_YUV420P16LE: return GO(L_I420, 16);
    default: return GMATB_ERR_UNSUPPORTED;
    }
#undef GO
}

int yuv2rgb_planar_launch(const GmatbImage *src, const GmatbImage *dst, const Mat9 &M, float norm,
                          const float shift[3], cudaStream_t st) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height) return GMATB_ERR_INVAL;
    if (dst->format != GMATB_FMT_RGBPF32LE && dst->format != GMATB_FMT_RGBAPF32LE) return GMATB_ERR_UNSUPPORTED;
    Img s, d;
    if (!to_img(src, &s, fmt_planes(src->format)) || !to_img(dst, &d, 3)) return GMATB_ERR_INVAL;
    const bool sparse = (M.m[1] == 0.f && M.m[8] == 0.f);
    const int vec = aligned16(s, fmt_planes(src->format)) && aligned16(d, 3);
    dim3 g = tile_grid(s.w, s.h, 4, src->batch), b(32, 8);        // a thread converts 8 columns x 4 rows
    float s0 = shift ? shift[0] : 0.f, s1 = shift ? shift[1] : 0.f, s2 = shift ? shift[2] : 0.f;
    if (src->format == GMATB_FMT_NV12) {
        if (sparse) yuv2rgb_planar_f32_kernel<L_NV12, true><<<g, b, 0, st>>>(s, d, M, norm, s0, s1, s2, vec);
        else        yuv2rgb_planar_f32_kernel<L_NV12, false><<<g, b, 0, st>>>(s, d, M, norm, s0, s1, s2, vec);
    } else if (src->format == GMATB_FMT_YUV420P) {
        if (sparse) yuv2rgb_planar_f32_kernel<L_I420, true><<<g, b, 0, st>>>(s, d, M, norm, s0, s1, s2, vec);
        else        yuv2rgb_planar_f32_kernel<L_I420, false><<<g, b, 0, st>>>(s, d, M, norm, s0, s1, s2, vec);
    } else return GMATB_ERR_UNSUPPORTED;
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

template <int SRC>
static int launch_rgb2yuv_src(int dfmt, dim3 g, cudaStream_t st, const Img &s, const Img &d, const Mat9 &M) {
    dim3 b(32, 8);
    const int np = (dfmt == GMATB_FMT_YUV420P) ? 3 : 2;
    const int vec = aligned16(s, 1) && aligned16(d, np);
    const dim3 g8 = tile_grid(s.w, s.h, 2 * r2y_tiles<SRC, 8>(), (int)g.z);
    switch (dfmt) {
    case GMATB_FMT_NV12:    rgb2yuv_kernel<SRC, L_NV12, 8><<<g8, b, 0, st>>>(s, d, M, vec); break;
    case GMATB_FMT_YUV420P: rgb2yuv_kernel<SRC, L_I420, 8><<<g8, b, 0, st>>>(s, d, M, vec); break;
    case GMATB_FMT_P010LE:
    case GMATB_FMT_P016LE:
        if (!srgb_is16(SRC)) return GMATB_ERR_UNSUPPORTED;
        rgb2yuv_kernel<SRC, L_NV12, 16><<<g, b, 0, st>>>(s, d, M, vec); break;
    default: return GMATB_ERR_UNSUPPORTED;
    }
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

int rgb2yuv_launch(const GmatbImage *src, const GmatbImage *dst, const Mat9 &M, cudaStream_t st) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height || src->width <= 0 || src->height <= 0)
        return GMATB_ERR_INVAL;
    Img s, d;
    if (!to_img(src, &s, 1) || !to_img(dst, &d, fmt_planes(dst->format))) return GMATB_ERR_INVAL;
    dim3 g = tile_grid(s.w, s.h, 2, src->batch);
    switch (src->format) {
    case GMATB_FMT_RGB24: return launch_rgb2yuv_src<S_RGB24>(dst->format, g, st, s, d, M);
    case GMATB_FMT_BGR24: return launch_rgb2yuv_src<S_BGR24>(dst->format, g, st, s, d, M);
    case GMATB_FMT_RGB0:
    case GMATB_FMT_RGBA:  return launch_rgb2yuv_src<S_RGBA>(dst->format, g, st, s, d, M);
    case GMATB_FMT_BGR0:
    case GMATB_FMT_BGRA:  return launch_rgb2yuv_src<S_BGRA>(dst->format, g, st, s, d, M);
    case GMATB_FMT_RGBA64LE: return launch_rgb2yuv_src<S_RGBA64>(dst->format, g, st, s, d, M);
    case GMATB_FMT_BGRA64LE: return launch_rgb2yuv_src<S_BGRA64>(dst->format, g, st, s, d, M);
    default: return GMATB_ERR_UNSUPPORTED;
    }
}

static bool yuv_desc(int fmt, int *layout, int *depth);

// The (source, destination) pairs the launchers above and below accept: gmatb_sws_create refuses the others, as
// the reference refuses them at sws_getContext time (ff_get_unscaled_swscale_cuda leaves convert_unscaled NULL,
// swscale_unscaled.c:2014-2054, and sws_init_context_cuda returns EINVAL).  kind: 0 yuv->rgb, 1 rgb->yuv, 2 yuv->yuv.
bool csc_pair_supported(int kind, int sf, int df) {
    int l, d;
    auto rgb_src = [](int f, bool *is16) {
        *is16 = (f == GMATB_FMT_RGBA64LE || f == GMATB_FMT_BGRA64LE);
        switch (f) {
        case GMATB_FMT_RGB24: case GMATB_FMT_BGR24: case GMATB_FMT_RGB0: case GMATB_FMT_RGBA: case GMATB_FMT_BGR0: case GMATB_FMT_BGRA:
        case GMATB_FMT_RGBA64LE: case GMATB_FMT_BGRA64LE: return true;
        default: return false;
        }
    };
    if (kind == 0) {
        if (!yuv_desc(sf, &l, &d)) return false;
        if (df == GMATB_FMT_RGBPF32LE) return sf == GMATB_FMT_NV12 || sf == GMATB_FMT_YUV420P;   // RGBAPF32LE: no alpha plane writer
        return dst_code(df) >= 0;
    }
    if (kind == 1) {
        bool is16;
        if (!rgb_src(sf, &is16)) return false;
        if (df == GMATB_FMT_NV12 || df == GMATB_FMT_YUV420P) return true;
        return (df == GMATB_FMT_P010LE || df == GMATB_FMT_P016LE) && is16;
    }
    return yuv_desc(sf, &l, &d) && yuv_desc(df, &l, &d);
}

static bool yuv_desc(int fmt, int *layout, int *depth) {
    switch (fmt) {
    case GMATB_FMT_NV12:        *layout = L_NV12; *depth = 8;  return true;
    case GMATB_FMT_YUV420P:     *layout = L_I420; *depth = 8;  return true;
    case GMATB_FMT_P010LE:      *layout = L_NV12; *depth = 10; return true;
    case GMATB_FMT_P016LE:      *layout = L_NV12; *depth = 16; return true;
    case GMATB_FMT_YUV420P10LE: *layout = L_I420; *depth = 10; return true;   // MSB-aligned, as the reference writes it
    case GMATB_FMT_YUV420P16LE: *layout = L_I420; *depth = 16; return true;
    default: return false;
    }
}

int yuv2yuv_launch(const GmatbImage *src, const GmatbImage *dst, cudaStream_t st) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height || src->width <= 0 || src->height <= 0)
        return GMATB_ERR_INVAL;
    int sl, sd, dl, dd;
    if (!yuv_desc(src->format, &sl, &sd) || !yuv_desc(dst->format, &dl, &dd)) return GMATB_ERR_UNSUPPORTED;
    Img s, d;
    if (!to_img(src, &s, fmt_planes(src->format)) || !to_img(dst, &d, fmt_planes(dst->format))) return GMATB_ERR_INVAL;
    dim3 g = tile_grid(s.w, s.h, 2, src->batch), b(32, 8);
    const int vec = aligned16(s, fmt_planes(src->format)) && aligned16(d, fmt_planes(dst->format));
    // same depth, whole 16-byte pieces, even height: the copy / (de)interleave kernel
    const int sbytes = sd == 8 ? 1 : 2;
    if (sd == dd && vec && (s.w * sbytes) % 16 == 0 && (s.h & 1) == 0) {
        dim3 gc((s.w * sbytes / 16 + 31) / 32, (s.h / 2 + 7) / 8, src->batch > 1 ? src->batch : 1);
#define KC(SL_, DL_) do { if (sbytes == 1) yuv2yuv_copy_kernel<SL_, DL_, 1><<<gc, b, 0, st>>>(s, d); else yuv2yuv_copy_kernel<SL_, DL_, 2><<<gc, b, 0, st>>>(s, d); } while (0)
        if (sl == L_NV12) { if (dl == L_NV12) KC(L_NV12, L_NV12); else KC(L_NV12, L_I420); }
        else              { if (dl == L_NV12) KC(L_I420, L_NV12); else KC(L_I420, L_I420); }
#undef KC
        count_launch();
        return set_cuda_error(cudaGetLastError());
    }
#define K(SL, SD, DL, DD) yuv2yuv_kernel<SL, SD, DL, DD><<<g, b, 0, st>>>(s, d, vec)
#define KD(SL, SD, DL) do { if (dd == 8) K(SL, SD, DL, 8); else if (dd == 10) K(SL, SD, DL, 10); else K(SL, SD, DL, 16); } while (0)
#define KL(SL, SD) do { if (dl == L_NV12) KD(SL, SD, L_NV12); else KD(SL, SD, L_I420); } while (0)
#define KS(SL) do { if (sd == 8) KL(SL, 8); else if (sd == 10) KL(SL, 10); else KL(SL, 16); } while (0)
    if (sl == L_NV12) KS(L_NV12); else KS(L_I420);
#undef K
#undef KD
#undef KL
#undef KS
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

int rgb24swap_launch(const GmatbImage *src, const GmatbImage *dst, cudaStream_t st) {
    if (!src || !dst || src->width != dst->width || src->height != dst->height || src->width <= 0 || src->height <= 0)
        return GMATB_ERR_INVAL;
    Img s, d;
    if (!to_img(src, &s, 1) || !to_img(dst, &d, 1)) return GMATB_ERR_INVAL;
    const int vec = aligned16(s, 1) && aligned16(d, 1);
    dim3 g = tile_grid(s.w, s.h, 1, src->batch), b(32, 8);
    rgb24_swap_kernel<<<g, b, 0, st>>>(s, d, vec);
    count_launch();
    return set_cuda_error(cudaGetLastError());
}

}  // namespace gmatb
