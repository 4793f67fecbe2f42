// rotate_linear.cuh -- bilinear rotate for word-aligned packed images, 4 destination pixels per thread.
//
// Same arithmetic as rotate_kernel / oracle orc_rotate (P-FILTERS): coordinates in double
// (dy*s and dy*c are shared by the 4 pixels of a thread, every product and sum individually
// rounded), the four taps fetched as aligned words + funnel shift, byte -> float through PRMT
// magic numbers, weights and the 4-term blend in the oracle's operation order but packed two
// pixels at a time on FFMA2, round-to-nearest-even by adding 1.5*2^23, 12 / 16 output bytes
// stored as words.  The general kernel needs ~125 instructions per pixel (it is issue bound);
// this one ~60.
#pragma once
#include "common.cuh"
#include "tma.cuh"

namespace gmatb {

#define GMATB_MAGIC15 12582912.0f   /* 1.5 * 2^23: RN(v + M) - M = rint(v), bits(v + M) - bits(M) = (int)rint(v) for |v| < 2^22 */

// the two horizontally adjacent pixels starting `off` bytes into the (word-aligned) frame as magic
// floats (2^23 + byte): aligned words + funnel shift
template <int BPP>
__device__ __forceinline__ void rot_fetch2(const uint8_t *ps, uint32_t off, float (&pa)[BPP], float (&pb)[BPP]) {
    const uint32_t *q = reinterpret_cast<const uint32_t *>(ps) + (off >> 2);
    if (BPP == 4) {
        const uint32_t v0 = __ldg(q), v1 = __ldg(q + 1);
        pa[0] = byte_magic<0>(v0); pa[1] = byte_magic<1>(v0); pa[2] = byte_magic<2>(v0); pa[BPP - 1] = byte_magic<3>(v0);
        pb[0] = byte_magic<0>(v1); pb[1] = byte_magic<1>(v1); pb[2] = byte_magic<2>(v1); pb[BPP - 1] = byte_magic<3>(v1);
    } else {
        const unsigned sh = (off & 3u) * 8;
        const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1);
        const uint32_t w2 = sh == 24 ? __ldg(q + 2) : 0u;      // bytes 6..8 past q only when misaligned by 3
        const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
        pa[0] = byte_magic<0>(v0); pa[1] = byte_magic<1>(v0); pa[2] = byte_magic<2>(v0);
        pb[0] = byte_magic<3>(v0); pb[1] = byte_magic<0>(v1); pb[2] = byte_magic<1>(v1);
    }
}

// One pixel in the oracle's scalar form: the rare cases the packed path does not cover -- a coordinate in
// (-0.5, 0) (weights leave [0,1]: the result needs the saturation) or the last source column (both
// horizontal taps are pixel W-1).  Out of line so that it costs the common path nothing.
template <int BPP>
__device__ __noinline__ uint32_t rot_slow_pixel(const uint8_t *ps, int pitch, int W, int H, float sx, float sy) {
    const int x1 = __float2int_rz(sx), y1 = __float2int_rz(sy);
    const int x2 = x1 + 1, y2 = y1 + 1;
    const int x2r = min(x2, W - 1), y2r = min(y2, H - 1);
    const float ax = __fsub_rn((float)x2, sx), bx = __fsub_rn(sx, (float)x1);
    const float ay = __fsub_rn((float)y2, sy), by = __fsub_rn(sy, (float)y1);
    const float w00 = __fmul_rn(ax, ay), w01 = __fmul_rn(bx, ay), w10 = __fmul_rn(ax, by), w11 = __fmul_rn(bx, by);
    const uint8_t *r0 = ps + (size_t)y1 * pitch, *r1 = ps + (size_t)y2r * pitch;
    uint32_t out = 0;
#pragma unroll
    for (int c = 0; c < BPP; c++) {
        float a = __fmul_rn((float)r0[(size_t)x1 * BPP + c], w00);
        a = __fmaf_rn((float)r0[(size_t)x2r * BPP + c], w01, a);
        a = __fmaf_rn((float)r1[(size_t)x1 * BPP + c], w10, a);
        a = __fmaf_rn((float)r1[(size_t)x2r * BPP + c], w11, a);
        out |= (uint32_t)min(max(__float2int_rn(a), 0), 255) << (8 * c);
    }
    return out;
}

// 4 consecutive destination pixels from (x0, y) of frame fz, taps gathered from global memory: every case
// (frame borders, pixels outside the source)
template <int BPP>
__device__ __forceinline__ void rotate4_global(const PImg &s, const PImg &d, const RotParams &R, int x0, int y, long long fz) {
    const uint8_t *ps = s.p + fz * s.bstride;
    const int W = s.w, H = s.h;
    const float Wf = (float)W, Hf = (float)H, Wm1 = (float)(W - 1), Hm1 = (float)(H - 1);
    const double dy = __dsub_rn((double)y, R.shy);
    const double dys = __dmul_rn(dy, R.s), dyc = __dmul_rn(dy, R.c);
    const double xd = (double)x0;

    float m[4][2][2][BPP];        // [pixel][row][tap][component], magic floats
    float sx[4], sy[4], fx1[4], fy1[4];
    bool valid[4], special[4];    // special: handled by rot_slow_pixel
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double dx = __dsub_rn(__dadd_rn(xd, (double)i), R.shx);
        sx[i] = (float)__dsub_rn(__dmul_rn(dx, R.c), dys);
        sy[i] = (float)__dadd_rn(__dmul_rn(dx, R.s), dyc);
        valid[i] = sx[i] > -0.5f && sx[i] < Wf && sy[i] > -0.5f && sy[i] < Hf;
        // truncation through the 2^23 magic number (the F2I / I2F units are 8x slower than the FP32 pipe);
        // pixels outside the frame get clamped coordinates: their loads stay in range, their result is discarded
        const float xm = __fadd_rz(fminf(fmaxf(sx[i], 0.0f), Wm1), GMATB_MAGIC);
        const float ym = __fadd_rz(fminf(fmaxf(sy[i], 0.0f), Hm1), GMATB_MAGIC);
        const int x1 = __float_as_int(xm) - 0x4B000000, y1 = __float_as_int(ym) - 0x4B000000;
        fx1[i] = __fadd_rn(xm, -GMATB_MAGIC); fy1[i] = __fadd_rn(ym, -GMATB_MAGIC);
        // the tap pair (x1, x1+1) of rows y1 and min(y1+1, H-1)
        special[i] = valid[i] && (sx[i] < 0.0f || sy[i] < 0.0f || x1 > W - 2);
        const uint32_t off0 = (uint32_t)y1 * (uint32_t)s.pitch + (uint32_t)min(x1, W - 2) * BPP;
        const uint32_t off1 = off0 + (y1 < H - 1 ? (uint32_t)s.pitch : 0u);
        rot_fetch2<BPP>(ps, off0, m[i][0][0], m[i][0][1]);
        rot_fetch2<BPP>(ps, off1, m[i][1][0], m[i][1][1]);
    }
    uint32_t ob[4][BPP];          // result bytes (low byte of each word)
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
        const f2 sx2 = pk(sx[i], sx[i + 1]), sy2 = pk(sy[i], sy[i + 1]);
        const f2 fx2 = pk(fx1[i], fx1[i + 1]), fy2 = pk(fy1[i], fy1[i + 1]);
        const f2 neg1 = bc(-1.0f);
        // ax = (x1 + 1) - sx, bx = sx - x1 (and the same in y): a - b as FFMA2(b, -1, a), exact product, one rounding
        const f2 ax = fma2(sx2, neg1, add2(fx2, bc(1.0f))), bx = fma2(fx2, neg1, sx2);
        const f2 ay = fma2(sy2, neg1, add2(fy2, bc(1.0f))), by = fma2(fy2, neg1, sy2);
        const f2 w00 = mul2(ax, ay), w01 = mul2(bx, ay), w10 = mul2(ax, by), w11 = mul2(bx, by);
#pragma unroll
        for (int c = 0; c < BPP; c++) {
            const f2 nm = bc(-GMATB_MAGIC);
            const f2 p00 = add2(pk(m[i][0][0][c], m[i + 1][0][0][c]), nm), p01 = add2(pk(m[i][0][1][c], m[i + 1][0][1][c]), nm);
            const f2 p10 = add2(pk(m[i][1][0][c], m[i + 1][1][0][c]), nm), p11 = add2(pk(m[i][1][1][c], m[i + 1][1][1][c]), nm);
            f2 a = mul2(p00, w00);
            a = fma2(p01, w01, a);
            a = fma2(p10, w10, a);
            a = fma2(p11, w11, a);
            int b0, b1;
            upki(add2(a, bc(GMATB_MAGIC15)), b0, b1);
            ob[i][c] = valid[i] ? (uint32_t)b0 : 0u;
            ob[i + 1][c] = valid[i + 1] ? (uint32_t)b1 : 0u;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (special[i]) {
            const uint32_t v = rot_slow_pixel<BPP>(ps, s.pitch, W, H, sx[i], sy[i]);
#pragma unroll
            for (int c = 0; c < BPP; c++) ob[i][c] = (v >> (8 * c)) & 0xFFu;
        }
    uint8_t *pd = d.p + fz * d.bstride + (size_t)y * d.pitch + (size_t)x0 * BPP;
    auto pack4 = [](uint32_t a, uint32_t b, uint32_t c, uint32_t e) {
        return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, e, 0x0040), 0x5410);
    };
    if (BPP == 3) {
        stg32(pd,     pack4(ob[0][0], ob[0][1], ob[0][2], ob[1][0]));
        stg32(pd + 4, pack4(ob[1][1], ob[1][2], ob[2][0], ob[2][1]));
        stg32(pd + 8, pack4(ob[2][2], ob[3][0], ob[3][1], ob[3][2]));
    } else {
        stg128(pd, make_uint4(pack4(ob[0][0], ob[0][1], ob[0][2], ob[0][BPP - 1]), pack4(ob[1][0], ob[1][1], ob[1][2], ob[1][BPP - 1]),
                              pack4(ob[2][0], ob[2][1], ob[2][2], ob[2][BPP - 1]), pack4(ob[3][0], ob[3][1], ob[3][2], ob[3][BPP - 1])));
    }
}

template <int BPP>
__global__ void __launch_bounds__(256) rotate_linear4_kernel(PImg s, PImg d, RotParams R) {
    // thread = 4 consecutive destination pixels; warp = 8 threads x 4 rows (32x4 pixels), CTA 32x32
    const int x0 = (blockIdx.x * 8 + threadIdx.x) * 4;
    const int y = blockIdx.y * 32 + threadIdx.y;
    if (x0 >= d.w || y >= d.h) return;              // host guarantees d.w % 4 == 0
    rotate4_global<BPP>(s, d, R, x0, y, blockIdx.z);
}

// ---------------------------------------------------------------------------------------------------------
// TMA-staged form.  A CTA produces a 32 x 32 destination tile.  The tile's footprint in the source is a rotated
// square; its bounding box (<= 48 x 48 pixels for any angle) is fetched by ONE cp.async.bulk.tensor into shared
// memory while the threads compute their coordinates, and the four taps of every pixel are then gathered from
// shared memory (the gather from global memory touched 11.7 sectors per request: L1-bound, profiles/r1c).
// Tiles whose footprint lies wholly inside the source take this path with no border logic at all (every pixel is
// valid, no clamps, no special pixels); tiles that touch the border of the source, or fall outside it, run
// rotate4_global.  Same arithmetic, same bytes.
// The source position is monotonic in x and in y (correctly rounded operations of fixed operands), so the extremes
// over the tile are at its corners, computed with the per-pixel formula itself.
#ifndef GMATB_ROT_MINB
#define GMATB_ROT_MINB 4
#endif
template <int BPP>
__device__ __forceinline__ void rot_fetch2_smem(const uint8_t *tile, uint32_t off, float (&pa)[BPP], float (&pb)[BPP]) {
    const uint32_t *q = reinterpret_cast<const uint32_t *>(tile) + (off >> 2);
    if (BPP == 4) {
        const uint32_t v0 = q[0], v1 = q[1];
        pa[0] = byte_magic<0>(v0); pa[1] = byte_magic<1>(v0); pa[2] = byte_magic<2>(v0); pa[BPP - 1] = byte_magic<3>(v0);
        pb[0] = byte_magic<0>(v1); pb[1] = byte_magic<1>(v1); pb[2] = byte_magic<2>(v1); pb[BPP - 1] = byte_magic<3>(v1);
    } else {
        const unsigned sh = (off & 3u) * 8;
        const uint32_t w0 = q[0], w1 = q[1], w2 = q[2];          // the row is padded: always readable
        const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh);
        pa[0] = byte_magic<0>(v0); pa[1] = byte_magic<1>(v0); pa[2] = byte_magic<2>(v0);
        pb[0] = byte_magic<3>(v0); pb[1] = byte_magic<0>(v1); pb[2] = byte_magic<1>(v1);
    }
}

template <int BPP>
__global__ void __launch_bounds__(256, GMATB_ROT_MINB) rotate_linear_tma_kernel(const __grid_constant__ CUtensorMap smap, PImg s, PImg d, RotParams R,
                                                                int box_x, int box_y) {
    extern __shared__ __align__(128) uint8_t rot_tile[];
    __shared__ uint64_t bar;
    __shared__ int tile_info[4];                                  // interior?, bx0, by0 and the corner minima as float bits
    __shared__ float tile_min[2];
    const int tx0 = blockIdx.x * 32, ty0 = blockIdx.y * 32;
    const int x0 = tx0 + threadIdx.x * 4, y = ty0 + threadIdx.y;
    const long long fz = blockIdx.z;
    const int W = s.w, H = s.h;
    // ONE thread maps the four corners of the tile (clipped to the destination) through the per-pixel formula, decides
    // whether the footprint lies inside the source and, if so, starts the TMA load (as 256 per-thread copies this
    // prologue was ~19 of the kernel's 117 instructions per pixel)
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        const int tx1 = min(tx0 + 31, d.w - 1), ty1 = min(ty0 + 31, d.h - 1);
        float cx[4], cy[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double dxk = __dsub_rn((double)((k & 1) ? tx1 : tx0), R.shx), dyk = __dsub_rn((double)((k & 2) ? ty1 : ty0), R.shy);
            cx[k] = (float)__dsub_rn(__dmul_rn(dxk, R.c), __dmul_rn(dyk, R.s));
            cy[k] = (float)__dadd_rn(__dmul_rn(dxk, R.s), __dmul_rn(dyk, R.c));
        }
        const float cminx = fminf(fminf(cx[0], cx[1]), fminf(cx[2], cx[3])), cmaxx = fmaxf(fmaxf(cx[0], cx[1]), fmaxf(cx[2], cx[3]));
        const float cminy = fminf(fminf(cy[0], cy[1]), fminf(cy[2], cy[3])), cmaxy = fmaxf(fmaxf(cy[0], cy[1]), fmaxf(cy[2], cy[3]));
        const int bx0_ = (int)cminx, by0_ = (int)cminy;             // used only when both are >= 0
        // the tile's first byte: TMA wants the innermost coordinate of a byte tensor on a 16-byte boundary (an unaligned one
        // raises an illegal-instruction error on B200: tools/tma_probe2.cu)
        const int xb_ = (bx0_ * BPP) & ~15;
        const bool in = cminx >= 0.0f && cminy >= 0.0f && cmaxx < (float)(W - 1) && cmaxy < (float)(H - 1) &&
                        ((int)cmaxx + 2) * BPP + 8 - xb_ <= box_x && (int)cmaxy + 2 - by0_ <= box_y;
        tile_info[0] = in; tile_info[1] = xb_; tile_info[2] = by0_;
        tile_min[0] = cminx; tile_min[1] = cminy;
        if (in) {
            mbar_init(&bar, 1);
            mbar_expect_tx(&bar, (uint32_t)(box_x * box_y));
            tma_load_3d(rot_tile, &smap, &bar, xb_, by0_, (int)fz);
        }
    }
    __syncthreads();
    const bool interior = tile_info[0] != 0;
    const int xb = tile_info[1], by0 = tile_info[2];
    const float cminx = tile_min[0], cminy = tile_min[1];
    if (!interior) {                                              // block-uniform
        if (x0 < d.w && y < d.h) rotate4_global<BPP>(s, d, R, x0, y, fz);
        return;
    }
    const bool active = x0 < d.w && y < d.h;
    // coordinates while the tile is in flight
    const double dy = __dsub_rn((double)y, R.shy);
    const double dys = __dmul_rn(dy, R.s), dyc = __dmul_rn(dy, R.c);
    const double xd = (double)x0;
    float sx[4], sy[4], fx1[4], fy1[4];
    uint32_t off[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const double dx = __dsub_rn(__dadd_rn(xd, (double)i), R.shx);
        sx[i] = (float)__dsub_rn(__dmul_rn(dx, R.c), dys);
        sy[i] = (float)__dadd_rn(__dmul_rn(dx, R.s), dyc);
        // inactive threads (partial tiles at the right / bottom of the destination) keep their reads inside the box
        const float sxc = active ? sx[i] : cminx, syc = active ? sy[i] : cminy;
        const float xm = __fadd_rz(sxc, GMATB_MAGIC), ym = __fadd_rz(syc, GMATB_MAGIC);
        const int x1 = __float_as_int(xm) - 0x4B000000, y1 = __float_as_int(ym) - 0x4B000000;
        fx1[i] = __fadd_rn(xm, -GMATB_MAGIC); fy1[i] = __fadd_rn(ym, -GMATB_MAGIC);
        sx[i] = sxc; sy[i] = syc;
        off[i] = (uint32_t)(y1 - by0) * (uint32_t)box_x + (uint32_t)(x1 * BPP - xb);
    }
    mbar_wait(&bar, 0);
    if (!active) return;
    uint32_t ob[4][BPP];
#pragma unroll
    for (int i = 0; i < 4; i += 2) {                              // two pixels at a time: 8 taps of BPP floats live, not 16
        float m[4][2][2][BPP];
#pragma unroll
        for (int k = i; k < i + 2; k++) {
            rot_fetch2_smem<BPP>(rot_tile, off[k], m[k][0][0], m[k][0][1]);
            rot_fetch2_smem<BPP>(rot_tile, off[k] + (uint32_t)box_x, m[k][1][0], m[k][1][1]);
        }
        const f2 sx2 = pk(sx[i], sx[i + 1]), sy2 = pk(sy[i], sy[i + 1]);
        const f2 fx2 = pk(fx1[i], fx1[i + 1]), fy2 = pk(fy1[i], fy1[i + 1]);
        const f2 neg1 = bc(-1.0f);
        const f2 ax = fma2(sx2, neg1, add2(fx2, bc(1.0f))), bx = fma2(fx2, neg1, sx2);
        const f2 ay = fma2(sy2, neg1, add2(fy2, bc(1.0f))), by = fma2(fy2, neg1, sy2);
        const f2 w00 = mul2(ax, ay), w01 = mul2(bx, ay), w10 = mul2(ax, by), w11 = mul2(bx, by);
#pragma unroll
        for (int c = 0; c < BPP; c++) {
            const f2 nm = bc(-GMATB_MAGIC);
            const f2 p00 = add2(pk(m[i][0][0][c], m[i + 1][0][0][c]), nm), p01 = add2(pk(m[i][0][1][c], m[i + 1][0][1][c]), nm);
            const f2 p10 = add2(pk(m[i][1][0][c], m[i + 1][1][0][c]), nm), p11 = add2(pk(m[i][1][1][c], m[i + 1][1][1][c]), nm);
            f2 a = mul2(p00, w00);
            a = fma2(p01, w01, a);
            a = fma2(p10, w10, a);
            a = fma2(p11, w11, a);
            int b0, b1;
            upki(add2(a, bc(GMATB_MAGIC15)), b0, b1);
            ob[i][c] = (uint32_t)b0; ob[i + 1][c] = (uint32_t)b1;
        }
    }
    uint8_t *pd = d.p + fz * d.bstride + (size_t)y * d.pitch + (size_t)x0 * BPP;
    auto pack4 = [](uint32_t a, uint32_t b, uint32_t c, uint32_t e) {
        return __byte_perm(__byte_perm(a, b, 0x0040), __byte_perm(c, e, 0x0040), 0x5410);
    };
    if (BPP == 3) {
        stg32(pd,     pack4(ob[0][0], ob[0][1], ob[0][2], ob[1][0]));
        stg32(pd + 4, pack4(ob[1][1], ob[1][2], ob[2][0], ob[2][1]));
        stg32(pd + 8, pack4(ob[2][2], ob[3][0], ob[3][1], ob[3][2]));
    } else {
        stg128(pd, make_uint4(pack4(ob[0][0], ob[0][1], ob[0][2], ob[0][BPP - 1]), pack4(ob[1][0], ob[1][1], ob[1][2], ob[1][BPP - 1]),
                              pack4(ob[2][0], ob[2][1], ob[2][2], ob[2][BPP - 1]), pack4(ob[3][0], ob[3][1], ob[3][2], ob[3][BPP - 1])));
    }
}

}  // namespace gmatb
