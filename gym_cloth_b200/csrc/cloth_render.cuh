// cloth_render.cuh - 224x224 image observations rendered on the GPU.
//
// Stands where gym-cloth exports an .obj with trimesh, starts `blender --background --python get_image_rep_279.py`,
// sleeps a second and reads the PNG back (gym_cloth/envs/cloth_env.py:212-330).  The scene is the one that script
// builds: the cloth mesh (two triangles per grid cell, cloth_env.py:226-231, smooth shaded, front and back coloured
// differently, get_image_rep_279.py:188-260), the white unit-square bed 0.05 below it (:143-156), for depth images
// a floor plane 0.25 below (:126-140, :455-462), one shadow-less constant-falloff point lamp (:467-469) and a
// pinhole camera 1.45 above the centre looking straight down, 40 mm lens on a 36 mm sensor (:114-122, :267-277).
// Depth images are the camera-space Z pass normalised over the frame (compute_depth, :398-411).
//
// Blender itself is not available in the build container, so this row has no pixel-level parity pin: the checker
// (oracle/render_oracle.py) is a numpy restatement of THIS design, and the tests pin geometry (which pixel shows
// which cloth point, which side is visible, depth ordering) rather than Blender's shading rounding.
//
// One CTA renders a band of rows of one environment: a depth|triangle-id key per sample lives in shared memory,
// triangles are scattered into it with atomicMin, then every pixel of the band is shaded once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <cstring>
#include "../../include/clothb200.h"

namespace clothb200 {
extern std::atomic<long long> g_launch_count;
void set_cuda_error(cudaError_t e, const char *where);

struct RenderDev {
    int W;                 // cloth grid width (points)
    int N, ntri, id_bits;
    int height, width, samples, band_rows, bands;
    float fpx;             // focal length in pixels
    float dnear, dfar;
    ClothB200Scene sc;
    ClothB200SceneEnv env;
};

struct Cam {
    float C[3];
    float R[9];            // camera-to-world rotation (columns = camera x, y, z axes in world coordinates)
};

__device__ __forceinline__ Cam make_cam(const RenderDev &D, int e) {
    Cam c;
    float deg[3];
    for (int i = 0; i < 3; i++) {
        c.C[i] = D.sc.cam_pos[i] + (D.env.cam_pos_offset ? D.env.cam_pos_offset[3 * e + i] : 0.f);
        deg[i] = D.env.cam_deg ? D.sc.cam_deg[i] + D.env.cam_deg[3 * e + i] : D.sc.cam_deg[i];
    }
    const float k = 0.017453292519943295f;
    const float cx = cosf(deg[0] * k), sx = sinf(deg[0] * k), cy = cosf(deg[1] * k), sy = sinf(deg[1] * k),
                cz = cosf(deg[2] * k), sz = sinf(deg[2] * k);
    // Blender 'XYZ' Euler: R = Rz * Ry * Rx
    c.R[0] = cz * cy; c.R[1] = cz * sy * sx - sz * cx; c.R[2] = cz * sy * cx + sz * sx;
    c.R[3] = sz * cy; c.R[4] = sz * sy * sx + cz * cx; c.R[5] = sz * sy * cx - cz * sx;
    c.R[6] = -sy;     c.R[7] = cy * sx;                c.R[8] = cy * cx;
    return c;
}

// world -> (screen x, screen y, 1/depth); depth = distance along the viewing axis (the camera looks down its -z)
__device__ __forceinline__ float3 project(const RenderDev &D, const Cam &c, float x, float y, float z) {
    const float px = x - c.C[0], py = y - c.C[1], pz = z - c.C[2];
    const float cxv = c.R[0] * px + c.R[3] * py + c.R[6] * pz;
    const float cyv = c.R[1] * px + c.R[4] * py + c.R[7] * pz;
    const float czv = c.R[2] * px + c.R[5] * py + c.R[8] * pz;
    const float d = -czv;
    const float w = 1.0f / d;
    return make_float3(0.5f * D.width + D.fpx * cxv * w, 0.5f * D.height - D.fpx * cyv * w, d > 0.f ? w : -1.0f);
}

// ray through screen position (sx, sy): world direction with unit depth along the viewing axis
__device__ __forceinline__ float3 ray_dir(const RenderDev &D, const Cam &c, float sx, float sy) {
    const float u = (sx - 0.5f * D.width) / D.fpx, v = (0.5f * D.height - sy) / D.fpx;
    return make_float3(c.R[0] * u + c.R[1] * v - c.R[2], c.R[3] * u + c.R[4] * v - c.R[5], c.R[6] * u + c.R[7] * v - c.R[8]);
}

__device__ __forceinline__ void tri_vertices(int W, int t, int &a, int &b, int &c) {
    const int cell = t >> 1, r = cell / (W - 1), col = cell - r * (W - 1), pp = r * W + col;
    if (t & 1) { a = pp + 1; b = pp + W; c = pp + W + 1; }       // cloth_env.py:230
    else       { a = pp;     b = pp + W; c = pp + 1; }           // cloth_env.py:229
}

__device__ __forceinline__ float srgb_oetf(float x) {
    x = fminf(fmaxf(x, 0.f), 1.f);
    return x <= 0.0031308f ? 12.92f * x : 1.055f * powf(x, 1.0f / 2.4f) - 0.055f;
}

__device__ __forceinline__ unsigned enc_float_ordered(float f) {   // monotone float -> uint for atomicMin/Max
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float dec_float_ordered(unsigned u) {
    const unsigned v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    float f;
#ifdef __CUDA_ARCH__
    f = __uint_as_float(v);
#else
    memcpy(&f, &v, 4);
#endif
    return f;
}

// DEPTH = false: colour image, BGR uint8 [n][H][W][3] (cv2.imread order, cloth_env.py:292)
// DEPTH = true : camera-space depth per pixel into zbuf [n][H][W] (1e10 = nothing hit) and the per-environment
//                min / max over hit pixels into minmax [n][2] (ordered-uint encoding), for the Normalize node
template <typename T, bool DEPTH>
__global__ void __launch_bounds__(256) render_kernel(RenderDev D, const T *__restrict__ pos_all, uint8_t *__restrict__ out,
                                                      float *__restrict__ zbuf, unsigned *__restrict__ minmax) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int e = blockIdx.x / D.bands, band = blockIdx.x - e * D.bands, tid = threadIdx.x, NT = blockDim.x;
    const int S = DEPTH ? 1 : D.samples;
    const int row0 = band * D.band_rows;
    const int rows = min(D.band_rows, D.height - row0);
    const int SW = D.width * S, SH = rows * S;                 // sample grid of this band
    float *Pw = (float *)smem_raw;                              // world xyz      [N][3]
    float *Ps = Pw + 3 * D.N;                                   // screen x, y, 1/depth [N][3]
    float *Nv = Ps + 3 * D.N;                                   // vertex normals [N][3]
    unsigned *key = (unsigned *)(Nv + 3 * D.N);                 // [SH][SW]
    __shared__ unsigned s_min, s_max;
    const Cam cam = make_cam(D, e);
    const T *pos = pos_all + (size_t)e * D.N * 4;
    for (int p = tid; p < D.N; p += NT) {
        const float x = (float)pos[4 * p], y = (float)pos[4 * p + 1], z = (float)pos[4 * p + 2];
        Pw[3 * p] = x; Pw[3 * p + 1] = y; Pw[3 * p + 2] = z;
        const float3 s = project(D, cam, x, y, z);
        Ps[3 * p] = s.x; Ps[3 * p + 1] = s.y; Ps[3 * p + 2] = s.z;
    }
    for (int i = tid; i < SH * SW; i += NT) key[i] = 0xffffffffu;
    if (tid == 0) { s_min = 0xffffffffu; s_max = 0u; }
    __syncthreads();
    const int W = D.W;
    if (!DEPTH) {
        // smooth shading: vertex normal = sum of the (area weighted) normals of the faces around the vertex, in a
        // fixed order so that the image is reproducible
        for (int p = tid; p < D.N; p += NT) {
            const int r = p / W, c = p - r * W;
            float nx = 0.f, ny = 0.f, nz = 0.f;
            auto face = [&](int a, int b, int cc) {
                const float ux = Pw[3 * b] - Pw[3 * a], uy = Pw[3 * b + 1] - Pw[3 * a + 1], uz = Pw[3 * b + 2] - Pw[3 * a + 2];
                const float vx = Pw[3 * cc] - Pw[3 * a], vy = Pw[3 * cc + 1] - Pw[3 * a + 1], vz = Pw[3 * cc + 2] - Pw[3 * a + 2];
                nx += uy * vz - uz * vy; ny += uz * vx - ux * vz; nz += ux * vy - uy * vx;
            };
            if (r < W - 1 && c < W - 1) face(p, p + W, p + 1);                                    // cell (r, c), first triangle
            if (r > 0 && c < W - 1) { face(p - W, p, p - W + 1); face(p - W + 1, p, p + 1); }       // cell (r-1, c), both
            if (c > 0 && r < W - 1) { face(p - 1, p - 1 + W, p); face(p, p - 1 + W, p + W); }       // cell (r, c-1), both
            if (r > 0 && c > 0) face(p - W, p - 1, p);                                            // cell (r-1, c-1), second
            const float inv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
            Nv[3 * p] = nx * inv; Nv[3 * p + 1] = ny * inv; Nv[3 * p + 2] = nz * inv;
        }
    }
    // ---- scatter: one thread per triangle, atomicMin of (quantised depth | triangle id) per covered sample ----
    const float inv_s = 1.0f / S;
    const float y_lo = (float)row0, y_hi = (float)(row0 + rows);
    const unsigned dq_max = (1u << (32 - D.id_bits)) - 1u;
    const float dscale = (float)dq_max / (D.dfar - D.dnear);
    for (int t = tid; t < D.ntri; t += NT) {
        int a, b, c;
        tri_vertices(W, t, a, b, c);
        const float ax = Ps[3 * a], ay = Ps[3 * a + 1], aw = Ps[3 * a + 2];
        const float bx = Ps[3 * b], by = Ps[3 * b + 1], bw = Ps[3 * b + 2];
        const float cx = Ps[3 * c], cy = Ps[3 * c + 1], cw = Ps[3 * c + 2];
        if (aw <= 0.f || bw <= 0.f || cw <= 0.f) continue;                         // behind the camera: never in an episode
        const float area = (bx - ax) * (cy - ay) - (by - ay) * (cx - ax);
        if (fabsf(area) < 1e-12f) continue;
        const float xmin = fminf(ax, fminf(bx, cx)), xmax = fmaxf(ax, fmaxf(bx, cx));
        const float ymin = fminf(ay, fminf(by, cy)), ymax = fmaxf(ay, fmaxf(by, cy));
        if (xmax < 0.f || xmin > (float)D.width || ymax < y_lo || ymin > y_hi) continue;
        // sample (i, j) of the band sits at screen ((j + 0.5) / S, row0 + (i + 0.5) / S)
        const int j0 = max(0, (int)floorf(xmin * S - 0.5f)), j1 = min(SW - 1, (int)ceilf(xmax * S - 0.5f));
        const int i0 = max(0, (int)floorf((ymin - y_lo) * S - 0.5f)), i1 = min(SH - 1, (int)ceilf((ymax - y_lo) * S - 0.5f));
        const float inv_area = 1.0f / area;
        for (int i = i0; i <= i1; i++) {
            const float sy = y_lo + (i + 0.5f) * inv_s;
            for (int j = j0; j <= j1; j++) {
                const float sx = (j + 0.5f) * inv_s;
                const float l0 = ((bx - sx) * (cy - sy) - (by - sy) * (cx - sx)) * inv_area;
                const float l1 = ((cx - sx) * (ay - sy) - (cy - sy) * (ax - sx)) * inv_area;
                const float l2 = 1.0f - l0 - l1;
                if (l0 < 0.f || l1 < 0.f || l2 < 0.f) continue;
                const float w = l0 * aw + l1 * bw + l2 * cw;
                const float d = 1.0f / w;
                if (!(d > D.dnear) || d >= D.dfar) continue;
                const unsigned q = (unsigned)((d - D.dnear) * dscale);
                atomicMin(&key[i * SW + j], (min(q, dq_max) << D.id_bits) | (unsigned)t);
            }
        }
    }
    __syncthreads();
    // ---- resolve: one thread per pixel ----
    const float kd = D.sc.diffuse_intensity * D.sc.lamp_energy;
    float front[3], back[3], bed[3];
    for (int i = 0; i < 3; i++) {
        front[i] = D.env.front ? D.env.front[3 * e + i] : D.sc.front[i];
        back[i] = D.env.back ? D.env.back[3 * e + i] : D.sc.back[i];
        bed[i] = D.env.bed ? D.env.bed[3 * e + i] : D.sc.bed[i];
    }
    if (D.env.swap_sides && D.env.swap_sides[e]) for (int i = 0; i < 3; i++) { const float t = front[i]; front[i] = back[i]; back[i] = t; }
    unsigned lmin = 0xffffffffu, lmax = 0u;
    for (int px = tid; px < rows * D.width; px += NT) {
        const int pi = px / D.width, pj = px - pi * D.width;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, depth = 1e10f;
        for (int si = 0; si < S; si++)
            for (int sj = 0; sj < S; sj++) {
                const int i = pi * S + si, j = pj * S + sj;
                const float sx = (j + 0.5f) * inv_s, sy = y_lo + (i + 0.5f) * inv_s;
                const float3 dir = ray_dir(D, cam, sx, sy);
                // planes below the cloth: bed (both modes) and floor (depth images only)
                float dplane = 1e10f; int which = 0;      // 0 background, 1 bed, 2 floor
                if (dir.z < 0.f) {
                    const float tb = (D.sc.bed_z - cam.C[2]) / dir.z;
                    const float bxw = cam.C[0] + tb * dir.x, byw = cam.C[1] + tb * dir.y;
                    if (tb > 0.f && bxw >= D.sc.bed_x0 && bxw <= D.sc.bed_x1 && byw >= D.sc.bed_y0 && byw <= D.sc.bed_y1) { dplane = tb; which = 1; }
                    else if (DEPTH) {
                        const float tf = (D.sc.floor_z - cam.C[2]) / dir.z;
                        const float fx = cam.C[0] + tf * dir.x, fy = cam.C[1] + tf * dir.y;
                        if (tf > 0.f && fx >= D.sc.floor_x0 && fx <= D.sc.floor_x1 && fy >= D.sc.floor_y0 && fy <= D.sc.floor_y1) { dplane = tf; which = 2; }
                    }
                }
                const unsigned k = key[i * SW + j];
                bool cloth = false;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, dcl = 1e10f;
                int a = 0, b = 0, c = 0;
                if (k != 0xffffffffu) {
                    tri_vertices(W, (int)(k & ((1u << D.id_bits) - 1u)), a, b, c);
                    const float ax = Ps[3 * a], ay = Ps[3 * a + 1], bx = Ps[3 * b], by = Ps[3 * b + 1], cx = Ps[3 * c], cy = Ps[3 * c + 1];
                    const float inv_area = 1.0f / ((bx - ax) * (cy - ay) - (by - ay) * (cx - ax));
                    l0 = ((bx - sx) * (cy - sy) - (by - sy) * (cx - sx)) * inv_area;
                    l1 = ((cx - sx) * (ay - sy) - (cy - sy) * (ax - sx)) * inv_area;
                    l2 = 1.0f - l0 - l1;
                    dcl = 1.0f / (l0 * Ps[3 * a + 2] + l1 * Ps[3 * b + 2] + l2 * Ps[3 * c + 2]);
                    cloth = dcl <= dplane;
                }
                if (DEPTH) { depth = cloth ? dcl : dplane; continue; }
                float col[3];
                if (cloth) {
                    // perspective-correct weights
                    float b0 = l0 * Ps[3 * a + 2] * dcl, b1 = l1 * Ps[3 * b + 2] * dcl, b2 = l2 * Ps[3 * c + 2] * dcl;
                    const float wx = b0 * Pw[3 * a] + b1 * Pw[3 * b] + b2 * Pw[3 * c], wy = b0 * Pw[3 * a + 1] + b1 * Pw[3 * b + 1] + b2 * Pw[3 * c + 1],
                                wz = b0 * Pw[3 * a + 2] + b1 * Pw[3 * b + 2] + b2 * Pw[3 * c + 2];
                    float nx = b0 * Nv[3 * a] + b1 * Nv[3 * b] + b2 * Nv[3 * c], ny = b0 * Nv[3 * a + 1] + b1 * Nv[3 * b + 1] + b2 * Nv[3 * c + 1],
                          nz = b0 * Nv[3 * a + 2] + b1 * Nv[3 * b + 2] + b2 * Nv[3 * c + 2];
                    // which side of the face looks at the camera (the Geometry node's front/back output, :218-222)
                    const float ux = Pw[3 * b] - Pw[3 * a], uy = Pw[3 * b + 1] - Pw[3 * a + 1], uz = Pw[3 * b + 2] - Pw[3 * a + 2];
                    const float vx = Pw[3 * c] - Pw[3 * a], vy = Pw[3 * c + 1] - Pw[3 * a + 1], vz = Pw[3 * c + 2] - Pw[3 * a + 2];
                    const float gx = uy * vz - uz * vy, gy = uz * vx - ux * vz, gz = ux * vy - uy * vx;
                    const float ex = cam.C[0] - wx, ey = cam.C[1] - wy, ez = cam.C[2] - wz;
                    const bool is_front = gx * ex + gy * ey + gz * ez >= 0.f;
                    if (nx * ex + ny * ey + nz * ez < 0.f) { nx = -nx; ny = -ny; nz = -nz; }   // shade the side that is seen
                    const float ninv = rsqrtf(fmaxf(nx * nx + ny * ny + nz * nz, 1e-30f));
                    const float lx = D.sc.lamp_pos[0] - wx, ly = D.sc.lamp_pos[1] - wy, lz = D.sc.lamp_pos[2] - wz;
                    const float linv = rsqrtf(lx * lx + ly * ly + lz * lz);
                    const float lam = fmaxf(0.f, (nx * lx + ny * ly + nz * lz) * ninv * linv) * kd;
                    for (int q = 0; q < 3; q++) col[q] = (is_front ? front[q] : back[q]) * lam;
                } else if (which == 1) {
                    const float wx = cam.C[0] + dplane * dir.x, wy = cam.C[1] + dplane * dir.y;
                    const float lx = D.sc.lamp_pos[0] - wx, ly = D.sc.lamp_pos[1] - wy, lz = D.sc.lamp_pos[2] - D.sc.bed_z;
                    const float lam = fmaxf(0.f, lz * rsqrtf(lx * lx + ly * ly + lz * lz)) * kd;
                    for (int q = 0; q < 3; q++) col[q] = bed[q] * lam;
                } else {
                    for (int q = 0; q < 3; q++) col[q] = D.sc.horizon;
                }
                acc0 += fminf(col[0], 1.f); acc1 += fminf(col[1], 1.f); acc2 += fminf(col[2], 1.f);
            }
        const size_t o = ((size_t)e * D.height + row0 + pi) * D.width + pj;
        if (DEPTH) {
            zbuf[o] = depth;
            if (depth < 1e9f) { const unsigned u = enc_float_ordered(depth); lmin = min(lmin, u); lmax = max(lmax, u); }
        } else {
            const float n = 1.0f / (S * S);
            out[3 * o + 2] = (uint8_t)(srgb_oetf(acc0 * n) * 255.f + 0.5f);     // R
            out[3 * o + 1] = (uint8_t)(srgb_oetf(acc1 * n) * 255.f + 0.5f);     // G
            out[3 * o + 0] = (uint8_t)(srgb_oetf(acc2 * n) * 255.f + 0.5f);     // B
        }
    }
    if (DEPTH) {
        atomicMin(&s_min, lmin); atomicMax(&s_max, lmax);
        __syncthreads();
        if (tid == 0) { atomicMin(&minmax[2 * e], s_min); atomicMax(&minmax[2 * e + 1], s_max); }
    }
}

static __global__ void depth_minmax_init_kernel(int n, unsigned *minmax) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { minmax[2 * i] = 0xffffffffu; minmax[2 * i + 1] = 0u; }
}

// Normalize node + the display transform of save_render (get_image_rep_279.py:398-411, 306-307): 8-bit grey
static __global__ void depth_normalise_kernel(int n, int hw, const float *__restrict__ zbuf, const unsigned *__restrict__ minmax, uint8_t *__restrict__ gray) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)n * hw) return;
    const int e = (int)(i / hw);
    const unsigned umin = minmax[2 * e], umax = minmax[2 * e + 1];
    const float z = zbuf[i];
    float v = 1.0f;
    if (z < 1e9f && umax >= umin) {
        const float zmin = dec_float_ordered(umin), zmax = dec_float_ordered(umax);
        v = zmax > zmin ? (z - zmin) / (zmax - zmin) : 0.f;
    }
    gray[i] = (uint8_t)(srgb_oetf(v) * 255.f + 0.5f);
}

inline size_t render_smem_bytes(int N, int width, int S, int band_rows) {
    return (size_t)9 * N * sizeof(float) + (size_t)band_rows * S * width * S * sizeof(unsigned);
}

template <typename T>
int render_t(const ClothB200Params *hp, const ClothB200Scene *sc, const ClothB200SceneEnv *env, int n_env, const T *pos, bool depth,
             uint8_t *out, float *zbuf, unsigned *minmax, cudaStream_t st) {
    if (!hp || !sc || !pos || n_env < 0) return CLOTHB200_ERR_ARG;
    if (depth ? (!zbuf || !minmax || !out) : !out) return CLOTHB200_ERR_ARG;
    if (n_env == 0) return CLOTHB200_OK;
    const int W = hp->num_width_points;
    if (W != hp->num_height_points || W < 2) return CLOTHB200_ERR_CONFIG;
    if (sc->height < 1 || sc->width < 1 || sc->height > 4096 || sc->width > 4096 || sc->samples < 1 || sc->samples > 4 ||
        !(sc->lens_mm > 0.f) || !(sc->sensor_mm > 0.f))
        return CLOTHB200_ERR_CONFIG;
    RenderDev D;
    D.W = W; D.N = W * W; D.ntri = 2 * (W - 1) * (W - 1);
    D.id_bits = 1; while ((1 << D.id_bits) <= D.ntri) D.id_bits++;
    D.height = sc->height; D.width = sc->width; D.samples = sc->samples;
    D.fpx = sc->lens_mm / sc->sensor_mm * (float)(sc->width > sc->height ? sc->width : sc->height);   // sensor fit AUTO
    D.dnear = 0.05f; D.dfar = 4.0f;
    D.sc = *sc;
    if (env) D.env = *env; else memset(&D.env, 0, sizeof(D.env));
    const int S = depth ? 1 : sc->samples;
    // two bands resident per SM; shrink the band if the cloth's vertex data is large
    size_t budget = 100 * 1024;
    const size_t fixed = (size_t)9 * D.N * sizeof(float);
    if (fixed + (size_t)S * S * sc->width * 4 > budget) budget = 200 * 1024;
    if (fixed + (size_t)S * S * sc->width * 4 > budget) return CLOTHB200_ERR_CONFIG;
    int rows = (int)((budget - fixed) / ((size_t)S * S * sc->width * 4));
    if (rows > sc->height) rows = sc->height;
    const int bands = (sc->height + rows - 1) / rows;
    rows = (sc->height + bands - 1) / bands;
    D.band_rows = rows; D.bands = bands;
    const size_t smem = render_smem_bytes(D.N, sc->width, S, rows);
    cudaError_t err;
    if (depth) {
        depth_minmax_init_kernel<<<(n_env + 255) / 256, 256, 0, st>>>(n_env, minmax);
        err = cudaFuncSetAttribute(render_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { set_cuda_error(err, "render attr"); return CLOTHB200_ERR_CUDA; }
        render_kernel<T, true><<<(unsigned)bands * (unsigned)n_env, 256, smem, st>>>(D, pos, nullptr, zbuf, minmax);
        const size_t tot = (size_t)n_env * sc->height * sc->width;
        depth_normalise_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(n_env, sc->height * sc->width, zbuf, minmax, out);
        g_launch_count += 3;
    } else {
        err = cudaFuncSetAttribute(render_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) { set_cuda_error(err, "render attr"); return CLOTHB200_ERR_CUDA; }
        render_kernel<T, false><<<(unsigned)bands * (unsigned)n_env, 256, smem, st>>>(D, pos, out, nullptr, nullptr);
        g_launch_count += 1;
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) { set_cuda_error(err, "render launch"); return CLOTHB200_ERR_CUDA; }
    return CLOTHB200_OK;
}

}  // namespace clothb200
