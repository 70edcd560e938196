// cloth_device.cuh - device-side building blocks of the batched cloth step (sm_100a).
//
// One CTA owns one cloth.  pos/prev of all N points live in shared memory for the whole action
// (staged in and out with TMA bulk copies); every phase of Cloth.update() (cloth.pyx:169-214) runs
// there.  REFERENCE_ORDER mode reproduces the sequential Gauss-Seidel semantics of the Cython loops
// exactly (bit-exact in the double instantiation) while still exposing parallelism:
//   - Hooke (cloth.pyx:221-237) is a Jacobi phase: per-point gather of its <=12 springs in global
//     spring order gives the same per-point summation order as the reference's scatter loop.
//   - self-collision (cloth.pyx:313-343) is sequential only inside one hash bucket.  All points are
//     first evaluated against the post-Verlet snapshot; a bucket whose snapshot has no hit is final;
//     in a bucket with hits the first hit point (lowest index) takes its snapshot correction and only
//     the points after it are replayed in index order by one warp (lanes = candidates).
//   - the 10 % stretch limit (cloth.pyx:258-296) modifies a spring's end points only if it is longer
//     than 1.1 x rest.  All springs are tested against the snapshot in parallel; flagged springs go
//     into a bitmask priority queue ordered by spring index and are replayed in order by one warp;
//     every applied correction re-tests the <=22 later springs incident to the moved points, so a
//     spring is processed exactly when the sequential loop would have found it stretched.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/clothb200.h"

namespace clothb200 {

// ------------------------------------------------------------------------------------------------
// scalar helpers
// ------------------------------------------------------------------------------------------------
template <typename T> struct V4;
template <> struct V4<float> { typedef float4 type; };
template <> struct V4<double> { typedef double4 type; };

__device__ __forceinline__ float4 mk4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
__device__ __forceinline__ double4 mk4(double x, double y, double z, double w) { return make_double4(x, y, z, w); }

__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float floor_t(float x) { return floorf(x); }
__device__ __forceinline__ double floor_t(double x) { return floor(x); }

// cloth.pyx:17-18 fastnorm: sqrt(x*x + y*y + z*z), left-to-right
template <typename T> __device__ __forceinline__ T norm3(T x, T y, T z) { return sqrt_t(x * x + y * y + z * z); }

template <typename T> struct DevParams {
    int W, H, N;
    int table_size, table_shift; // hash table: power of two >= 1.6 N
    int ev_words;                // ceil(6N/32) words of the stretched-spring queue
    int n_levels;                // grab_top z levels (gripper.pyx:31-41)
    int max_actions;
    T mg;                        // mass*gravity (cloth.pyx:179)
    T kk_struct, kk_bend;        // ks*1.0, ks*0.2 (cloth.pyx:225-232)
    T dsdm, damp;                // (dt*dt)/mass, 1-damping/100 (cloth.pyx:240-241)
    T cell_w, cell_h, cell_t;    // 3dx, 3dy, max (cloth.pyx:308-310)
    T thresh;                    // 2*thickness (cloth.pyx:317)
    T sim_steps;                 // simulation_steps as scalar (cloth.pyx:338-340)
    T min_z, fric1, surf_off;    // minimum_z, 1-plane_friction, 1e-4 (cloth.pyx:345-370)
    T tear_thresh;
    T rest_k[6];                 // rest lengths of the flat grid per spring kind k (used when rest == NULL)
    double grip_radius, thickness, gripper_height;
    double iu, iur, igr, ir;     // iters_up, iters_up_rest, iters_grip_rest, iters_rest
};

template <typename T> struct StepArgs {
    T *pos, *prev;
    const T *rest;
    long long rest_env_stride;
    const ClothB200Plan *plans;  // NULL => plain update mode (n_updates x Cloth.update())
    const double *grab_xy;       // grab-only mode (clothb200_grab_top_*)
    double grab_radius;
    int n_updates;
    int mode;                    // 0 step, 1 update_n, 2 grab only, 3 measure only
    int initialize;
    int n_env;
    int32_t *flags, *sim_steps, *n_grabbed;
    uint32_t *grab_mask;
    double *coverage, *variance_inv;
    T *obs;
    double *prev_coverage;
    int32_t *num_steps, *num_sim_steps;
    double *reward;
    int32_t *done;
    const double *iters_up_env;
    const int32_t *env_order;
    long long *prof;             // optional [n_env][16] cycle / event counters (debug)
};

enum { KMODE_STEP = 0, KMODE_UPDATE = 1, KMODE_GRAB = 2, KMODE_MEASURE = 3 };

// ------------------------------------------------------------------------------------------------
// TMA bulk copy + mbarrier (PTX).  cp.async.bulk needs 16-byte aligned addresses and sizes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// The per-CTA cloth.  NT threads; WC = compile-time grid width (0 = runtime).
// ------------------------------------------------------------------------------------------------
#define CLOTH_KEY_EMPTY 0x7fffffff
#define CLOTH_FIRST_NONE 0x7ffffffe

template <typename T, int NT, int WC, bool REST_TABLE> struct ClothCTA {
    typedef typename V4<T>::type P4;
    static constexpr int NWARPS = NT / 32;

    const DevParams<T> &P;
    const int W, H, N;
    const int tid, lane, warp;
    P4 *pos, *prev;          // [N]
    int *tkey;               // [TS]   hash keys, later the bucket's first hit point
    uint32_t *tinfo;         // [TS]   count | (end offset << 16)
    uint16_t *pslot;         // [N]    bucket slot of each point (bit 15: this point allocates the bucket)
    uint16_t *lstA, *lstB;   // [N]    bucket member lists: unordered / index-ordered; lstA later = fix-up work list
    uint32_t *ev;            // [ev_words] stretched-spring queue
    int *misc;               // [16] counters/flags: 0 total, 1 nwork, 2 tear, 3 bad, 4.. scratch
    const T *rest;           // rest table of this env (REST_TABLE)
    long long pacc[16];      // thread 0: cycles per phase + event counters when profiling
    long long plast;
    bool prof_on;

    __device__ ClothCTA(const DevParams<T> &P_, unsigned char *smem, const T *rest_)
        : P(P_), W(WC ? WC : P_.W), H(WC ? WC : P_.H), N(WC ? WC * WC : P_.N), tid(threadIdx.x), lane(threadIdx.x & 31),
          warp(threadIdx.x >> 5), rest(rest_), plast(0), prof_on(false) {
#pragma unroll
        for (int i = 0; i < 16; i++) pacc[i] = 0;
        size_t o = 0;
        pos = reinterpret_cast<P4 *>(smem + o); o += sizeof(P4) * (size_t)N;
        prev = reinterpret_cast<P4 *>(smem + o); o += sizeof(P4) * (size_t)N;
        tkey = reinterpret_cast<int *>(smem + o); o += 4 * (size_t)P.table_size;
        tinfo = reinterpret_cast<uint32_t *>(smem + o); o += 4 * (size_t)P.table_size;
        ev = reinterpret_cast<uint32_t *>(smem + o); o += 4 * (size_t)((P.ev_words + 3) & ~3);
        misc = reinterpret_cast<int *>(smem + o); o += 4 * 16;
        pslot = reinterpret_cast<uint16_t *>(smem + o); o += 2 * (size_t)((N + 7) & ~7);
        lstA = reinterpret_cast<uint16_t *>(smem + o); o += 2 * (size_t)((N + 7) & ~7);
        lstB = reinterpret_cast<uint16_t *>(smem + o); o += 2 * (size_t)((N + 7) & ~7);
    }
    static __host__ __device__ size_t smem_bytes(int N, int table_size, int ev_words) {
        return sizeof(P4) * (size_t)N * 2 + 8 * (size_t)table_size + 4 * (size_t)((ev_words + 3) & ~3) + 64 +
               3 * 2 * (size_t)((N + 7) & ~7) + 16 /* mbarrier */;
    }

    __device__ __forceinline__ void sync() {
        if (NT == 32) __syncwarp(); else __syncthreads();
    }

    // ---- spring topology (cloth.pyx:135-146).  Spring slot s = q*6 + k: k-th spring created by point q. ----
    // offset from q back to ptA for kind k
    __device__ __forceinline__ int koff(int k) const {
        switch (k) {
            case 0: return W;      // STRUCTURAL (r-1,c)
            case 1: return 1;      // STRUCTURAL (r,c-1)
            case 2: return W + 1;  // SHEARING   (r-1,c-1)
            case 3: return W - 1;  // SHEARING   (r-1,c+1)
            case 4: return 2 * W;  // BENDING    (r-2,c)
            default: return 2;     // BENDING    (r,c-2)
        }
    }
    // does point (r,c) create spring kind k?
    __device__ __forceinline__ bool kvalid(int r, int c, int k) const {
        switch (k) {
            case 0: return r > 0;
            case 1: return c > 0;
            case 2: return r > 0 && c > 0;
            case 3: return r > 0 && c + 1 < W;
            case 4: return r > 1;
            default: return c > 1;
        }
    }
    __device__ __forceinline__ T rest_of(int q, int k) const {
        if (REST_TABLE) return __ldg(rest + q * 6 + k);
        return P.rest_k[k];
    }
    __device__ __forceinline__ T kk_of(int k) const { return k >= 4 ? P.kk_bend : P.kk_struct; }

    // ---- Hooke (cloth.pyx:221-237) gathered per point + Verlet (cloth.pyx:239-256) ----
    // The new position is parked in prev[p].xyz (prev is private to the owner thread); commit_verlet()
    // swaps it in after the barrier, when no thread reads old positions any more.
    __device__ __forceinline__ void spring_force(const P4 &Pa, const P4 &Pb, T kk, T rst, T &f0, T &f1, T &f2) {
        // l2_norm_ab = fastnorm(pb-pa); force_mg = ks*kc*(l-rest)/l; force_on_a = force_mg*(pb-pa)
        T d0 = Pb.x - Pa.x, d1 = Pb.y - Pa.y, d2 = Pb.z - Pa.z;
        T l = norm3(d0, d1, d2);
        if (l == T(0)) { misc[3] = 1; f0 = f1 = f2 = T(0); return; }  // reference: ZeroDivisionError
        T fm = kk * (l - rst) / l;
        f0 = fm * d0; f1 = fm * d1; f2 = fm * d2;
    }

    __device__ void hooke_verlet() {
        for (int p = tid; p < N; p += NT) {
            const P4 Pp = pos[p];
            if (Pp.w != T(0)) continue;  // pinned: Verlet skips it, its force is never used
            const int r = p / W, c = p - r * W;
            // _reset_gravity: f = 0 then f += (0,0,mg)
            T fx = T(0) + T(0), fy = T(0) + T(0), fz = T(0) + P.mg;
            T a0, a1, a2;
            // springs created by p (p is ptB): ptB.add_force(-F)
#pragma unroll
            for (int k = 0; k < 6; k++) {
                if (kvalid(r, c, k)) {
                    const int a = p - koff(k);
                    spring_force(pos[a], Pp, kk_of(k), rest_of(p, k), a0, a1, a2);
                    fx = fx + (-a0); fy = fy + (-a1); fz = fz + (-a2);
                }
            }
            // springs created by later points in which p is ptA, in creation order:
            // q = p+1 (k=1), p+2 (k=5), p+W-1 (k=3), p+W (k=0), p+W+1 (k=2), p+2W (k=4)
            if (c + 1 < W) { spring_force(Pp, pos[p + 1], P.kk_struct, rest_of(p + 1, 1), a0, a1, a2); fx = fx + a0; fy = fy + a1; fz = fz + a2; }
            if (c + 2 < W) { spring_force(Pp, pos[p + 2], P.kk_bend, rest_of(p + 2, 5), a0, a1, a2); fx = fx + a0; fy = fy + a1; fz = fz + a2; }
            if (r + 1 < H) {
                if (c > 0) { spring_force(Pp, pos[p + W - 1], P.kk_struct, rest_of(p + W - 1, 3), a0, a1, a2); fx = fx + a0; fy = fy + a1; fz = fz + a2; }
                { spring_force(Pp, pos[p + W], P.kk_struct, rest_of(p + W, 0), a0, a1, a2); fx = fx + a0; fy = fy + a1; fz = fz + a2; }
                if (c + 1 < W) { spring_force(Pp, pos[p + W + 1], P.kk_struct, rest_of(p + W + 1, 2), a0, a1, a2); fx = fx + a0; fy = fy + a1; fz = fz + a2; }
            }
            if (r + 2 < H) { spring_force(Pp, pos[p + 2 * W], P.kk_bend, rest_of(p + 2 * W, 4), a0, a1, a2); fx = fx + a0; fy = fy + a1; fz = fz + a2; }
            // Verlet: new = x + damp*(x-px) + f*dsdm
            const P4 Q = prev[p];
            T nx = Pp.x + (P.damp * (Pp.x - Q.x)) + (fx * P.dsdm);
            T ny = Pp.y + (P.damp * (Pp.y - Q.y)) + (fy * P.dsdm);
            T nz = Pp.z + (P.damp * (Pp.z - Q.z)) + (fz * P.dsdm);
            prev[p] = mk4(nx, ny, nz, Q.w);
        }
    }

    // ---- commit Verlet + build_spatial_map pass 1 (cloth.pyx:298-311): find/insert bucket, count ----
    __device__ __forceinline__ int cell(T v, T w) {
        T f = floor_t(v / w);
        if (!(f > T(-1048576) && f < T(1048576))) { misc[3] = 1; f = T(0); }  // NaN/huge: reference raises
        return (int)f;
    }
    __device__ void commit_and_hash() {
        for (int p = tid; p < N; p += NT) {
            P4 Pp = pos[p];
            if (Pp.w == T(0)) {
                const P4 Nw = prev[p];
                prev[p] = mk4(Pp.x, Pp.y, Pp.z, Nw.w);
                Pp = mk4(Nw.x, Nw.y, Nw.z, T(0));
                pos[p] = Pp;
            }
            const int key = 961 * cell(Pp.x, P.cell_w) + 31 * cell(Pp.y, P.cell_h) + cell(Pp.z, P.cell_t);
            uint32_t slot = ((uint32_t)key * 2654435761u) >> P.table_shift;
            const uint32_t msk = (uint32_t)P.table_size - 1;
            for (;;) {
                int old = *reinterpret_cast<volatile int *>(&tkey[slot]);
                if (old == key) break;
                if (old == CLOTH_KEY_EMPTY) {
                    old = atomicCAS(&tkey[slot], CLOTH_KEY_EMPTY, key);
                    if (old == CLOTH_KEY_EMPTY || old == key) break;
                }
                slot = (slot + 1) & msk;
            }
            const uint32_t r = atomicAdd(&tinfo[slot], 1u);
            pslot[p] = (uint16_t)(slot | (r == 0 ? 0x8000u : 0u));
        }
    }
    // pass 2: the first arriver of each bucket reserves its range in the member list
    __device__ void alloc_buckets() {
        for (int p = tid; p < N; p += NT) {
            const uint32_t s = pslot[p];
            if (s & 0x8000u) {
                const uint32_t slot = s & 0x7fffu;
                const uint32_t cnt = tinfo[slot];
                const uint32_t off = (uint32_t)atomicAdd(&misc[0], (int)cnt);
                tinfo[slot] = cnt | (off << 16);
                tkey[slot] = CLOTH_FIRST_NONE;
            }
        }
    }
    // pass 3: unordered scatter (high half of tinfo runs from off to off+cnt)
    __device__ void scatter_members() {
        for (int p = tid; p < N; p += NT) {
            const uint32_t slot = pslot[p] & 0x7fffu;
            const uint32_t idx = atomicAdd(&tinfo[slot], 0x10000u) >> 16;
            lstA[idx] = (uint16_t)p;
        }
    }
    // pass 4: rank inside the bucket -> index-ordered list (the dict value order of cloth.pyx:301-305)
    __device__ void order_members() {
        for (int p = tid; p < N; p += NT) {
            const uint32_t info = tinfo[pslot[p] & 0x7fffu];
            const int cnt = info & 0xffffu, start = (int)(info >> 16) - cnt;
            int rk = 0;
            for (int j = 0; j < cnt; j++) rk += (lstA[start + j] < p) ? 1 : 0;
            lstB[start + rk] = (uint16_t)p;
        }
    }

    // ---- self_collide (cloth.pyx:313-343) of point p against its bucket, current positions ----
    __device__ __forceinline__ int collide_point(int p, const P4 &Pp, int start, int cnt, T &cx, T &cy, T &cz) {
        T t0 = T(0), t1 = T(0), t2 = T(0);
        int n = 0;
        for (int j = 0; j < cnt; j++) {
            const int q = lstB[start + j];
            if (q == p) continue;
            const P4 Pq = pos[q];
            T d0 = Pp.x - Pq.x, d1 = Pp.y - Pq.y, d2 = Pp.z - Pq.z;
            T d = norm3(d0, d1, d2);
            if (d <= P.thresh) {
                if (d == T(0)) { misc[3] = 1; continue; }
                T factor = (P.thresh - d) / d;
                t0 += d0 * factor; t1 += d1 * factor; t2 += d2 * factor;
                n += 1;
            }
        }
        if (n) {
            T nf = (T)n;
            cx = t0 / nf / P.sim_steps; cy = t1 / nf / P.sim_steps; cz = t2 / nf / P.sim_steps;
        }
        return n;
    }
    // snapshot evaluation of every point: only records, per bucket, the first point that has a hit
    __device__ void collide_snapshot() {
        for (int p = tid; p < N; p += NT) {
            const P4 Pp = pos[p];
            if (Pp.w != T(0)) continue;
            const uint32_t slot = pslot[p] & 0x7fffu;
            const uint32_t info = tinfo[slot];
            const int cnt = info & 0xffffu, start = (int)(info >> 16) - cnt;
            if (cnt < 2) continue;
            T cx, cy, cz;
            if (collide_point(p, Pp, start, cnt, cx, cy, cz)) atomicMin(&tkey[slot], p);
        }
    }
    // _handle_plane_collision (cloth.pyx:345-370)
    __device__ __forceinline__ void plane_point(int p) {
        const P4 Pp = pos[p];
        if (Pp.w != T(0) || Pp.z >= P.min_z) return;
        const P4 Q = prev[p];
        T t = (P.min_z - Q.z) * T(1.0);
        T tx = Q.x + t * T(-0.0), ty = Q.y + t * T(-0.0), tz = Q.z + t * T(-1.0);
        T gx = tx + P.surf_off * T(0.0), gy = ty + P.surf_off * T(0.0), gz = tz + P.surf_off * T(1.0);
        T cx = gx - Q.x, cy = gy - Q.y, cz = gz - Q.z;
        pos[p] = mk4(Q.x + cx * P.fric1, Q.y + cy * P.fric1, Q.z + cz * P.fric1, Pp.w);
    }
    // first hit points take their snapshot correction and queue their bucket for the ordered replay;
    // points of buckets without any hit are final and get their plane collision here.
    __device__ void collide_first_and_plane() {
        for (int p = tid; p < N; p += NT) {
            const P4 Pp = pos[p];
            if (Pp.w != T(0)) continue;
            const uint32_t slot = pslot[p] & 0x7fffu;
            const uint32_t info = tinfo[slot];
            const int cnt = info & 0xffffu, start = (int)(info >> 16) - cnt;
            const int first = tkey[slot];
            // A bucket with a hit is finished by collide_replay(): its members must keep their
            // pre-plane positions while the first hit point re-reads them below.
            const bool replay = (first != CLOTH_FIRST_NONE);
            if (first == p) {
                T cx, cy, cz;
                if (collide_point(p, Pp, start, cnt, cx, cy, cz)) pos[p] = mk4(Pp.x + cx, Pp.y + cy, Pp.z + cz, Pp.w);
                lstA[atomicAdd(&misc[1], 1)] = (uint16_t)slot;
            }
            if (!replay) plane_point(p);
        }
    }
    // ordered replay of one bucket by one warp: points after the first hit, in index order;
    // lanes evaluate candidates, contributions are summed in candidate order.
    __device__ void collide_replay() {
        const int nwork = misc[1];
        for (int wi = warp; wi < nwork; wi += NWARPS) {
            const uint32_t slot = lstA[wi];
            const uint32_t info = tinfo[slot];
            const int cnt = info & 0xffffu, start = (int)(info >> 16) - cnt;
            const int first = tkey[slot];
            int j0 = 0;
            for (int base = 0; base < cnt; base += 32) {
                const int j = base + lane;
                const unsigned m = __ballot_sync(0xffffffffu, j < cnt && lstB[start + j] == first);
                if (m) { j0 = base + __ffs(m) - 1; break; }
            }
            for (int j = j0 + 1; j < cnt; j++) {
                const int p = lstB[start + j];
                const P4 Pp = pos[p];
                if (Pp.w != T(0)) continue;
                T t0 = T(0), t1 = T(0), t2 = T(0);
                int n = 0;
                for (int base = 0; base < cnt; base += 32) {
                    const int cj = base + lane;
                    const bool valid = cj < cnt && cj != j;
                    T c0 = T(0), c1 = T(0), c2 = T(0);
                    bool hit = false;
                    if (valid) {
                        const P4 Pq = pos[lstB[start + cj]];
                        T d0 = Pp.x - Pq.x, d1 = Pp.y - Pq.y, d2 = Pp.z - Pq.z;
                        T d = norm3(d0, d1, d2);
                        if (d <= P.thresh) {
                            if (d == T(0)) misc[3] = 1;
                            else {
                                T factor = (P.thresh - d) / d;
                                c0 = d0 * factor; c1 = d1 * factor; c2 = d2 * factor;
                                hit = true;
                            }
                        }
                    }
                    unsigned m = __ballot_sync(0xffffffffu, hit);
                    n += __popc(m);
                    while (m) {
                        const int l = __ffs(m) - 1;
                        m &= m - 1;
                        t0 += __shfl_sync(0xffffffffu, c0, l);
                        t1 += __shfl_sync(0xffffffffu, c1, l);
                        t2 += __shfl_sync(0xffffffffu, c2, l);
                    }
                }
                if (n && lane == 0) {
                    T nf = (T)n;
                    T cx = t0 / nf / P.sim_steps, cy = t1 / nf / P.sim_steps, cz = t2 / nf / P.sim_steps;
                    pos[p] = mk4(Pp.x + cx, Pp.y + cy, Pp.z + cz, Pp.w);
                }
                __syncwarp();
            }
            for (int base = 0; base < cnt; base += 32)
                if (base + lane < cnt) plane_point(lstB[start + base + lane]);
            __syncwarp();
        }
    }

    // ---- _limit_spring_changes (cloth.pyx:258-296) ----
    __device__ __forceinline__ bool spring_flagged(const P4 &Pa, const P4 &Pb, T rst) {
        if (Pa.w != T(0) && Pb.w != T(0)) return false;
        T l = norm3(Pa.x - Pb.x, Pa.y - Pb.y, Pa.z - Pb.z);
        return (l > rst * P.tear_thresh) || (l > (rst * T(1.1)));
    }
    // snapshot test of all springs; also clears the hash table for the next substep
    __device__ void limit_snapshot() {
        for (int j = tid; j < P.table_size; j += NT) { tkey[j] = CLOTH_KEY_EMPTY; tinfo[j] = 0u; }
        for (int p = tid; p < N; p += NT) {
            const P4 Pb = pos[p];
            const int r = p / W, c = p - r * W;
            uint32_t bits = 0;
#pragma unroll
            for (int k = 0; k < 6; k++)
                if (kvalid(r, c, k) && spring_flagged(pos[p - koff(k)], Pb, rest_of(p, k))) bits |= 1u << k;
            if (bits) {
                const int s = p * 6;
                const int sh = s & 31;
                const uint32_t lo = bits << sh, hi = sh > 26 ? bits >> (32 - sh) : 0u;
                if (lo) atomicOr(&ev[s >> 5], lo);
                if (hi) atomicOr(&ev[(s >> 5) + 1], hi);
            }
        }
    }
    // ordered replay by warp 0 (all lanes execute the spring update redundantly, lane 0 stores)
    __device__ void limit_replay() {
        if (warp != 0) return;
        const int nw = P.ev_words;
        int cursor = 0;
        for (;;) {
            // pop the lowest flagged spring slot >= cursor
            int s = 0x7fffffff;
            for (int base = cursor >> 5; base < nw; base += 32) {
                const int wi = base + lane;
                uint32_t word = wi < nw ? ev[wi] : 0u;
                if (wi == (cursor >> 5)) word &= ~((1u << (cursor & 31)) - 1u);
                const int cand = word ? (wi << 5) + __ffs(word) - 1 : 0x7fffffff;
                s = __reduce_min_sync(0xffffffffu, cand);
                if (s != 0x7fffffff) break;
            }
            if (s == 0x7fffffff) break;
            cursor = s + 1;
            if (prof_on && tid == 0) pacc[12] += 1;
            const int q = s / 6, k = s - q * 6;
            const int a = q - koff(k);
            P4 Pa = pos[a], Pb = pos[q];
            const bool pa = Pa.w != T(0), pb = Pb.w != T(0);
            if (pa && pb) continue;
            const T rst = rest_of(q, k);
            T l = norm3(Pa.x - Pb.x, Pa.y - Pb.y, Pa.z - Pb.z);
            if (l > rst * P.tear_thresh) misc[2] = 1;
            if (!(l > (rst * T(1.1)))) continue;
            T d0 = (Pa.x - Pb.x) / l, d1 = (Pa.y - Pb.y) / l, d2 = (Pa.z - Pb.z) / l;
            T extra = l - rst * T(1.1);
            if (prof_on && tid == 0) pacc[13] += 1;
            if (pa) {
                Pb = mk4(Pb.x + d0 * extra, Pb.y + d1 * extra, Pb.z + d2 * extra, Pb.w);
            } else if (pb) {
                Pa = mk4(Pa.x - d0 * extra, Pa.y - d1 * extra, Pa.z - d2 * extra, Pa.w);
            } else {
                T ed = extra * T(0.5);
                Pa = mk4(Pa.x - d0 * ed, Pa.y - d1 * ed, Pa.z - d2 * ed, Pa.w);
                Pb = mk4(Pb.x + d0 * ed, Pb.y + d1 * ed, Pb.z + d2 * ed, Pb.w);
            }
            if (lane == 0) { if (!pa) pos[a] = Pa; if (!pb) pos[q] = Pb; }
            __syncwarp();
            // re-test the later springs incident to the moved points: lanes 0-11 -> a, 12-23 -> q
            if (lane < 24) {
                const int x = lane < 12 ? a : q;
                const bool moved = lane < 12 ? !pa : !pb;
                const int i = lane < 12 ? lane : lane - 12;
                int qq, kk;
                if (i < 6) { qq = x; kk = i; }
                else {
                    // springs created by later points with x as ptA
                    switch (i) {
                        case 6: qq = x + 1; kk = 1; break;
                        case 7: qq = x + 2; kk = 5; break;
                        case 8: qq = x + W - 1; kk = 3; break;
                        case 9: qq = x + W; kk = 0; break;
                        case 10: qq = x + W + 1; kk = 2; break;
                        default: qq = x + 2 * W; kk = 4; break;
                    }
                }
                const int s2 = qq * 6 + kk;
                if (moved && qq < N && s2 > s) {
                    const int r2 = qq / W, c2 = qq - r2 * W;
                    const int aa = qq - koff(kk);
                    // for i >= 6 the spring must really connect x: kvalid() rules out row wrap-around
                    if (kvalid(r2, c2, kk) && (i < 6 || aa == x)) {
                        if (spring_flagged(pos[aa], pos[qq], rest_of(qq, kk))) atomicOr(&ev[s2 >> 5], 1u << (s2 & 31));
                    }
                }
            }
            __syncwarp();
        }
        // the queue is left clean for the next substep
        for (int j = lane; j < nw; j += 32) ev[j] = 0u;
    }

    __device__ __forceinline__ void ptick(int k) {
        if (prof_on && tid == 0) { const long long t = clock64(); pacc[k] += t - plast; plast = t; }
    }
    // ---- one Cloth.update() (cloth.pyx:169-214), reference order ----
    __device__ void update_reference_order() {
        if (prof_on && tid == 0) plast = clock64();
        hooke_verlet();            sync(); ptick(0);
        commit_and_hash();         sync(); ptick(1);
        alloc_buckets();           sync(); ptick(2);
        scatter_members();         sync(); ptick(3);
        order_members();           sync(); ptick(4);
        collide_snapshot();        sync(); ptick(5);
        collide_first_and_plane(); sync(); ptick(6);
        if (prof_on && tid == 0) pacc[11] += misc[1];
        collide_replay();          sync(); ptick(7);
        limit_snapshot();          sync(); ptick(8);
        limit_replay();
        if (tid == 0) { misc[0] = 0; misc[1] = 0; }
        sync(); ptick(9);
    }

    // ---- Gripper (gripper.pyx) ----
    // adjust (gripper.pyx:55-66): a point listed m times in grabbed_pts is moved m times
    __device__ void gripper_adjust(T dx, T dy, T dz) {
        for (int p = tid; p < N; p += NT) {
            P4 Q = prev[p];
            if (Q.w > T(0)) {
                P4 Pp = pos[p];
                const int m = (int)Q.w;
                for (int i = 0; i < m; i++) {
                    Q = mk4(Pp.x, Pp.y, Pp.z, Q.w);
                    Pp = mk4(dx + Pp.x, dy + Pp.y, dz + Pp.z, Pp.w);
                }
                pos[p] = Pp; prev[p] = Q;
            }
        }
    }
    // release (gripper.pyx:68-73)
    __device__ void gripper_release() {
        for (int p = tid; p < N; p += NT) {
            P4 Q = prev[p];
            if (Q.w > T(0)) {
                P4 Pp = pos[p];
                Pp.w = T(0); Q.w = T(0);
                pos[p] = Pp; prev[p] = Q;
            }
        }
    }
    // grab_top (gripper.pyx:23-42) in double.  Returns len(grabbed_pts) afterwards (all threads).
    // scratch: doubles in the (idle) hash table area.
    __device__ int grab_top(double x, double y, double radius) {
        double *levels = reinterpret_cast<double *>(tkey);
        const int nlev = P.n_levels;
        if (tid == 0) {
            double curZ = P.gripper_height;
            for (int i = 0; i < nlev; i++) { levels[i] = curZ; curZ -= P.thickness; }
            misc[4] = 0x7fffffff; misc[5] = 0;
        }
        sync();
        const double band = 2 * P.thickness;
        int mylev_min = 0x7fffffff;
        for (int p = tid; p < N; p += NT) {
            const P4 Pp = pos[p];
            const double px = (double)Pp.x, py = (double)Pp.y, pz = (double)Pp.z;
            if ((px - x) * (px - x) + (py - y) * (py - y) < radius) {
                for (int i = 0; i < nlev && i < mylev_min; i++)
                    if (fabs(pz - levels[i]) < band) { mylev_min = i; break; }
            }
        }
        if (mylev_min != 0x7fffffff) atomicMin(&misc[4], mylev_min);
        sync();
        const int lev = misc[4];
        int cnt = 0;
        for (int p = tid; p < N; p += NT) {
            P4 Pp = pos[p];
            P4 Q = prev[p];
            bool sel = false;
            if (lev != 0x7fffffff) {
                const double px = (double)Pp.x, py = (double)Pp.y, pz = (double)Pp.z;
                sel = ((px - x) * (px - x) + (py - y) * (py - y) < radius) && (fabs(pz - levels[lev]) < band);
            }
            if (sel) {
                Pp.w = T(1); Q.w = Q.w + T(1);
                pos[p] = Pp; prev[p] = Q;
            }
            cnt += (int)Q.w;
        }
        if (cnt) atomicAdd(&misc[5], cnt);
        sync();
        const int total = misc[5];
        sync();
        // restore the hash table area
        for (int j = tid; j < P.table_size; j += NT) { tkey[j] = CLOTH_KEY_EMPTY; tinfo[j] = 0u; }
        sync();
        return total;
    }

    // bit p set <=> point p is in gripper.grabbed_pts
    __device__ void write_grab_mask(uint32_t *out) {
        const int nwords = (N + 31) >> 5;
        for (int wi = tid; wi < nwords; wi += NT) {
            uint32_t word = 0;
            for (int b = 0; b < 32; b++) { const int p = wi * 32 + b; if (p < N && prev[p].w > T(0)) word |= 1u << b; }
            out[wi] = word;
        }
    }

    // ---- coverage: area of the convex hull of the clipped (x,y) (cloth_env.py:1086-1098) ----
    __device__ __forceinline__ void clipped_xy(int p, double &x, double &y) const {
        const P4 Pp = pos[p];
        double vx = (double)Pp.x, vy = (double)Pp.y;
        vx = (0.0 > vx) ? 0.0 : vx; vx = (1.0 < vx) ? 1.0 : vx;   // min(max(p.x,0),1)
        vy = (0.0 > vy) ? 0.0 : vy; vy = (1.0 < vy) ? 1.0 : vy;
        x = vx; y = vy;
    }
    __device__ __forceinline__ bool xy_less(int a, int b) const {
        if (a >= N) return false;   // padding sorts last
        if (b >= N) return true;
        double ax, ay, bx, by;
        clipped_xy(a, ax, ay); clipped_xy(b, bx, by);
        return (ax < bx) || (ax == bx && ay < by);
    }
    // Andrew monotone chain + shoelace in double, same operation order as oracle_hull_area.
    // Uses the hash table area as scratch (the table is rebuilt from scratch every substep).
    __device__ double hull_area() {
        uint16_t *idx = reinterpret_cast<uint16_t *>(tkey);   // [np2]
        uint16_t *hull = reinterpret_cast<uint16_t *>(tinfo); // [2N]
        int np2 = 1; while (np2 < N) np2 <<= 1;
        for (int i = tid; i < np2; i += NT) idx[i] = (uint16_t)i;
        sync();
        for (int k = 2; k <= np2; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < np2; i += NT) {
                    const int l = i ^ j;
                    if (l > i) {
                        const int a = idx[i], b = idx[l];
                        const bool up = (i & k) == 0;
                        if (up ? xy_less(b, a) : xy_less(a, b)) { idx[i] = (uint16_t)b; idx[l] = (uint16_t)a; }
                    }
                }
                sync();
            }
        }
        double area = 0.0;
        if (tid == 0) {
            int kk = 0;
            double ox, oy, ax, ay, bx, by;
            for (int i = 0; i < N; i++) {
                clipped_xy(idx[i], bx, by);
                while (kk >= 2) {
                    clipped_xy(hull[kk - 2], ox, oy); clipped_xy(hull[kk - 1], ax, ay);
                    if ((ax - ox) * (by - oy) - (ay - oy) * (bx - ox) <= 0) kk--; else break;
                }
                hull[kk++] = idx[i];
            }
            for (int i = N - 2, t = kk + 1; i >= 0; i--) {
                clipped_xy(idx[i], bx, by);
                while (kk >= t) {
                    clipped_xy(hull[kk - 2], ox, oy); clipped_xy(hull[kk - 1], ax, ay);
                    if ((ax - ox) * (by - oy) - (ay - oy) * (bx - ox) <= 0) kk--; else break;
                }
                hull[kk++] = idx[i];
            }
            kk--;
            double a2 = 0.0, hx, hy;
            clipped_xy(hull[0], hx, hy);
            for (int i = 0; i < kk; i++) {
                clipped_xy(hull[i], ax, ay);
                clipped_xy(hull[(i + 1) % kk], bx, by);
                a2 += (ax - hx) * (by - hy) - (bx - hx) * (ay - hy);
            }
            area = (kk >= 3) ? 0.5 * fabs(a2) : 0.0;
        }
        sync();
        for (int j = tid; j < P.table_size; j += NT) { tkey[j] = CLOTH_KEY_EMPTY; tinfo[j] = 0u; }
        sync();
        return area;  // valid in thread 0
    }

    // block-wide double sum (result valid in all threads); scratch = first words of tinfo, which is
    // idle (all zero) outside update() and is zeroed again before returning
    __device__ double block_sum(double v) {
        double *buf = reinterpret_cast<double *>(tinfo);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        sync();
        if (lane == 0) buf[warp] = v;
        sync();
        double s = 0.0;
        for (int w = 0; w < NWARPS; w++) s += buf[w];
        sync();
        if (tid < 2 * NWARPS) tinfo[tid] = 0u;
        sync();
        return s;
    }
    // _compute_variance (cloth_env.py:1075-1084)
    __device__ double variance_inv() {
        double s = 0.0;
        for (int p = tid; p < N; p += NT) s += (double)pos[p].z;
        const double mean = block_sum(s) / N;
        double v = 0.0;
        for (int p = tid; p < N; p += NT) { const double d = (double)pos[p].z - mean; v += d * d; }
        const double var = block_sum(v) / N;
        return (var < 0.000001) ? 1000.0 : 0.001 / var;
    }
    // _out_of_bounds (cloth_env.py:1020-1045): bounds (1,1,1), slack 0.25
    __device__ bool out_of_bounds() {
        int bad = 0;
        for (int p = tid; p < N; p += NT) {
            const P4 Pp = pos[p];
            const double x = (double)Pp.x, y = (double)Pp.y, z = (double)Pp.z;
            bad |= (x >= 1 + 0.25) || (x < -0.25) || (y >= 1 + 0.25) || (y < -0.25) || (z >= 1) || (z < 0);
        }
        if (NT == 32) return __any_sync(0xffffffffu, bad) != 0;
        return __syncthreads_or(bad) != 0;
    }
};

}  // namespace clothb200
