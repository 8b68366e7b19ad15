// Stage 2 -- DoG extrema scan + sub-pixel refinement, hand-written for sm_100a.
//
// Replaces find_extrema_in_dog / is_extremum / ModeFunctions / solve
// (reference src/popsift/s_extrema.cu:22-558, s_solve.h:25-86).  One launch covers every
// octave and level of an image (the reference launches once per octave): a flat list of
// 32x8-pixel tiles over all (octave, level) pairs; each warp scans one 32-pixel row segment.
// The contrast pre-test uses one coalesced read of the centre plane; the 26 neighbours are only
// fetched (through L1/L2) by the few lanes that pass it.  Survivors are refined in registers and
// appended with one warp-ballot + one atomicAdd per warp.
//
// Parity contract: with bit-identical DoG planes, the accepted set and (x, y, lpos) of every
// extremum are bit-identical to the reference's.  The floating-point pattern below is the one in
// the reference's sm_100 SASS (cuobjdump of oracle/_ref/libpopsift_ref.so):
//   detK = fma(a1,a2, -(b1*b2));  det = fma(i02,det2, fma(i00,det0, i01*det1));  rsd = __frcp_rn(det)
//   d    = rows of (adj*rsd) . b as fma chains starting from fma(.,b.x,0)
//   contr= v + 0.5*fma(dz,Dz, fma(dy,Dy, dx*Dx));  det2 = fma(DDx,DDy, -(DXx*DXx))
#include "ps_internal.h"

namespace psb {

namespace {

struct ScanParams {
    // tile lists: tile_begin[o] = first tile index of octave o (tiles of all L levels of an octave
    // are contiguous); tiles_x[o] = tiles per row
    int tile_begin[kMaxOctaves + 1];
    int tiles_x[kMaxOctaves];
    int tiles_per_level[kMaxOctaves];
};

struct DogView {
    const float* base;
    int w, h, pitch, nplanes;
    size_t plane;
    __device__ __forceinline__ float at(int x, int y, int z) const
    {   // clamp addressing in x, y and layer, like the reference's clamp textures
        x = min(max(x, 0), w - 1);
        y = min(max(y, 0), h - 1);
        z = min(max(z, 0), nplanes - 1);
        return __ldg(base + z * plane + (size_t)y * pitch + x);
    }
};

__device__ __forceinline__ bool strict_extremum(const DogView& d, int x, int y, int z, float val)
{
    bool gt = true, lt = true;
#pragma unroll
    for (int dz = -1; dz <= 1; ++dz)
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dx == 0 && dy == 0 && dz == 0) continue;
                const float f = d.at(x + dx, y + dy, z + dz);
                gt = gt && (val > f);
                lt = lt && (val < f);
            }
    return gt || lt;
}

__device__ __forceinline__ bool solve3(float i00, float i01, float i02, float i11, float i12, float i22,
                                       float& bx, float& by, float& bz)
{
    const float det0 = __fmaf_rn(i11, i22, -__fmul_rn(i12, i12));
    const float det1 = __fmaf_rn(i12, i02, -__fmul_rn(i01, i22));
    const float det2 = __fmaf_rn(i01, i12, -__fmul_rn(i11, i02));
    const float det3 = __fmaf_rn(i00, i22, -__fmul_rn(i02, i02));
    const float det4 = __fmaf_rn(i01, i02, -__fmul_rn(i00, i12));
    const float det5 = __fmaf_rn(i00, i11, -__fmul_rn(i01, i01));
    float det = __fmul_rn(i01, det1);
    det = __fmaf_rn(i00, det0, det);
    det = __fmaf_rn(i02, det2, det);
    if (det == 0.0f) return false;
    const float rsd = __frcp_rn(det);
    const float a00 = __fmul_rn(det0, rsd), a10 = __fmul_rn(det1, rsd), a20 = __fmul_rn(det2, rsd);
    const float a11 = __fmul_rn(det3, rsd), a12 = __fmul_rn(det4, rsd), a22 = __fmul_rn(det5, rsd);
    const float X = bx, Y = by, Z = bz;
    bx = __fmaf_rn(a20, Z, __fmaf_rn(a10, Y, __fmaf_rn(a00, X, 0.0f)));
    by = __fmaf_rn(a12, Z, __fmaf_rn(a11, Y, __fmaf_rn(a10, X, 0.0f)));
    bz = __fmaf_rn(a22, Z, __fmaf_rn(a12, Y, __fmaf_rn(a20, X, 0.0f)));
    return true;
}

template <int MODE>
__device__ bool refine(const DogView& dv, const Consts& k, int x, int y, int level, int maxlevel, float val,
                       InitialExtremum& out)
{
    const int width = dv.w, height = dv.h;
    float Dx = 0.f, Dy = 0.f, Dz = 0.f, DDx = 0.f, DDy = 0.f, DXx = 0.f;
    float dx = 0.f, dy = 0.f, dz = 0.f;
    int nx = x, ny = y, nz = level;
    int iter = 0;
    do {
        ++iter;
        const float x2y1z1 = dv.at(nx + 1, ny, nz), x0y1z1 = dv.at(nx - 1, ny, nz);
        const float x1y2z1 = dv.at(nx, ny + 1, nz), x1y0z1 = dv.at(nx, ny - 1, nz);
        const float x1y1z2 = dv.at(nx, ny, nz + 1), x1y1z0 = dv.at(nx, ny, nz - 1);
        Dx = __fmul_rn(__fsub_rn(x2y1z1, x0y1z1), 0.5f);
        Dy = __fmul_rn(__fsub_rn(x1y2z1, x1y0z1), 0.5f);
        Dz = __fmul_rn(__fsub_rn(x1y1z2, x1y1z0), 0.5f);
        const float c2 = __fmul_rn(dv.at(nx, ny, nz), 2.0f);
        DDx = __fsub_rn(__fadd_rn(x2y1z1, x0y1z1), c2);
        DDy = __fsub_rn(__fadd_rn(x1y2z1, x1y0z1), c2);
        const float DDz = __fsub_rn(__fadd_rn(x1y1z2, x1y1z0), c2);
        const float x0y0z1 = dv.at(nx - 1, ny - 1, nz), x0y1z0 = dv.at(nx - 1, ny, nz - 1);
        const float x0y1z2 = dv.at(nx - 1, ny, nz + 1), x0y2z1 = dv.at(nx - 1, ny + 1, nz);
        const float x1y0z0 = dv.at(nx, ny - 1, nz - 1), x1y0z2 = dv.at(nx, ny - 1, nz + 1);
        const float x1y2z0 = dv.at(nx, ny + 1, nz - 1), x1y2z2 = dv.at(nx, ny + 1, nz + 1);
        const float x2y0z1 = dv.at(nx + 1, ny - 1, nz), x2y1z0 = dv.at(nx + 1, ny, nz - 1);
        const float x2y1z2 = dv.at(nx + 1, ny, nz + 1), x2y2z1 = dv.at(nx + 1, ny + 1, nz);
        DXx = __fmul_rn(__fsub_rn(__fsub_rn(__fadd_rn(x2y2z1, x0y0z1), x0y2z1), x2y0z1), 0.25f);
        const float DXy = __fmul_rn(__fsub_rn(__fsub_rn(__fadd_rn(x2y1z2, x0y1z0), x0y1z2), x2y1z0), 0.25f);
        const float DXz = __fmul_rn(__fsub_rn(__fsub_rn(__fadd_rn(x1y2z2, x1y0z0), x1y2z0), x1y0z2), 0.25f);

        float bx = -Dx, by = -Dy, bz = -Dz;
        if (!solve3(DDx, DXx, DXy, DDy, DXz, DDz, bx, by, bz)) { dx = dy = dz = 0.0f; break; }
        dx = bx; dy = by; dz = bz;

        int retval;
        if (MODE == PS_MODE_OPENCV) {
            if (fabsf(dx) < 0.5f && fabsf(dy) < 0.5f && fabsf(dz) < 0.5f) retval = 1;
            else {
                nx += (int)roundf(dx); ny += (int)roundf(dy); nz += (int)roundf(dz);
                retval = (nx < 5 || nx >= width - 5 || ny < 5 || ny >= height - 5 || nz < 1 || nz > maxlevel - 2) ? -1 : 0;
            }
        } else if (iter == 5) {
            retval = 0;
        } else {
            const int tx = ((dx >= 0.6f && nx < width - 2) ? 1 : 0) + ((dx <= -0.6f && nx > 1) ? -1 : 0);
            const int ty = ((dy >= 0.6f && ny < height - 2) ? 1 : 0) + ((dy <= -0.6f && ny > 1) ? -1 : 0);
            int tz = 0;
            if (MODE == PS_MODE_POPSIFT)
                tz = ((dz >= 0.6f && nz < maxlevel - 1) ? 1 : 0) + ((dz <= -0.6f && nz > 1) ? -1 : 0);
            if (tx == 0 && ty == 0 && tz == 0) retval = 1;
            else { nx += tx; ny += ty; nz += tz; retval = 0; }
        }
        if (retval == -1) return false;
        if (retval == 1) break;
    } while (iter < 5);

    if (MODE == PS_MODE_OPENCV) { if (iter >= 5) return false; }
    else if (dx >= 1.5f || dy >= 1.5f || dz >= 1.5f) return false;

    const float xn = __fadd_rn((float)nx, dx), yn = __fadd_rn((float)ny, dy), sn = __fadd_rn((float)nz, dz);
    if (MODE != PS_MODE_OPENCV)
        if (xn < 0.0f || xn > (float)width - 1.0f || yn < 0.0f || yn > (float)height - 1.0f ||
            sn < 0.0f || sn > (float)maxlevel) return false;

    const float t = __fmaf_rn(dz, Dz, __fmaf_rn(dy, Dy, __fmul_rn(dx, Dx)));
    const float contr = __fadd_rn(val, __fmul_rn(t, 0.5f));
    const float tr = __fadd_rn(DDx, DDy);
    const float det = __fmaf_rn(DDx, DDy, -__fmul_rn(DXx, DXx));
    const float edgeval = __fdiv_rn(__fmul_rn(tr, tr), det);
    if (!(det > 0.0f)) return false;
    if (fabsf(contr) < __fmul_rn(k.threshold, 2.0f)) return false;
    const float el1 = __fadd_rn(k.edge_limit, 1.0f);
    if (edgeval >= __fdiv_rn(__fmul_rn(el1, el1), k.edge_limit)) return false;

    out.xpos = xn;
    out.ypos = yn;
    out.lpos = (int)roundf(sn);
    out.sigma = __fmul_rn(k.sigma0, powf(k.sigma_k, sn));
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(256)
find_extrema_kernel(PyramidView pyr, Consts k, ScanParams sp, InitialExtremum* __restrict__ iext, Counters* ct)
{
    // which octave does this tile belong to?
    const int tile = blockIdx.x;
    int o = 0;
    while (o + 1 < pyr.num_octaves && tile >= sp.tile_begin[o + 1]) ++o;
    const OctaveView& ov = pyr.oct[o];
    const int local = tile - sp.tile_begin[o];
    const int lvl_idx = local / sp.tiles_per_level[o];
    const int t2 = local - lvl_idx * sp.tiles_per_level[o];
    const int ty = t2 / sp.tiles_x[o];
    const int tx = t2 - ty * sp.tiles_x[o];
    const int level = lvl_idx + 1;
    const int maxlevel = pyr.levels + 2;      // reference passes _levels-1 (s_extrema.cu:597)

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int x = tx * 32 + lane + 1;
    const int y = ty * 8 + warp + 1;

    DogView dv;
    dv.base = ov.dog; dv.w = ov.w; dv.h = ov.h; dv.pitch = ov.pitch; dv.plane = ov.plane;
    dv.nplanes = pyr.levels + 2;

    bool found = false;
    InitialExtremum e;
    e.xpos = e.ypos = e.sigma = 0.f; e.lpos = 0;
    // border voxels can never be strict extrema under clamp addressing (reference quirk)
    bool inside = (x <= ov.w - 2) && (y <= ov.h - 2);
    if (MODE == PS_MODE_OPENCV) inside = inside && !(x < 5 || y < 5 || x >= ov.w - 5 || y >= ov.h - 5);
    if (inside) {
        const float val = __ldg(ov.dog + (size_t)level * ov.plane + (size_t)y * ov.pitch + x);
        bool ok;
        if (MODE == PS_MODE_OPENCV) ok = fabsf(val) >= floorf(k.threshold);
        else if (MODE == PS_MODE_VLFEAT) ok = fabsf(val) >= __fmul_rn(__fmul_rn(0.8f, 2.0f), k.threshold);
        else ok = fabsf(val) >= __fmul_rn(1.6f, k.threshold);
        if (ok && strict_extremum(dv, x, y, level, val))
            found = refine<MODE>(dv, k, x, y, level, maxlevel, val, e);
    }
    const unsigned mask = __ballot_sync(0xffffffffu, found);
    if (mask == 0) return;
    int base = 0;
    if (lane == 0) base = atomicAdd(&ct->ext_ct[o], __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    const int idx = base + __popc(mask & ((1u << lane) - 1u));
    if (found && idx < k.max_extrema) iext[(size_t)o * k.max_extrema + idx] = e;
}

} // namespace

int launch_find_extrema(const PyramidView& pyr, const Consts& k, InitialExtremum* iext, Counters* ct, cudaStream_t st)
{
    ScanParams sp;
    int total = 0;
    for (int o = 0; o < pyr.num_octaves; ++o) {
        const int tx = (pyr.oct[o].w - 2 + 31) / 32;   // x in [1, w-2]
        const int ty = (pyr.oct[o].h - 2 + 7) / 8;
        sp.tile_begin[o] = total;
        sp.tiles_x[o] = tx > 0 ? tx : 1;
        sp.tiles_per_level[o] = (tx > 0 && ty > 0) ? tx * ty : 0;
        if (sp.tiles_per_level[o] == 0) { sp.tiles_per_level[o] = 1; sp.tile_begin[o] = total; total += 0; sp.tile_begin[o + 1] = total; continue; }
        total += sp.tiles_per_level[o] * pyr.levels;
        sp.tile_begin[o + 1] = total;
    }
    for (int o = pyr.num_octaves; o < kMaxOctaves; ++o) { sp.tile_begin[o + 1] = total; sp.tiles_x[o] = 1; sp.tiles_per_level[o] = 1; }
    if (total == 0) return 0;
    switch (k.sift_mode) {
        case PS_MODE_OPENCV: find_extrema_kernel<PS_MODE_OPENCV><<<total, 256, 0, st>>>(pyr, k, sp, iext, ct); break;
        case PS_MODE_VLFEAT: find_extrema_kernel<PS_MODE_VLFEAT><<<total, 256, 0, st>>>(pyr, k, sp, iext, ct); break;
        default:             find_extrema_kernel<PS_MODE_POPSIFT><<<total, 256, 0, st>>>(pyr, k, sp, iext, ct); break;
    }
    return 1;
}

} // namespace psb
