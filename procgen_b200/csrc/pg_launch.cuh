// Kernels, launch wrappers and the per-game dispatch table. Included by pg_runtime.cu (host runtime +
// C ABI) and by one translation unit per game (games_tu/tu_<game>.cu), so the 16 games compile in
// parallel; a game's kernels are instantiated only in its own unit.
#pragma once
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pg_kernels.cuh"

#ifndef PG_HOSTSIM
#include <cuda_runtime.h>
#endif

namespace pg {

// ================================================================= errors (cpp-utils.cpp:8-20)
static inline void pg_fatal(const char *fmt, ...) {
    fprintf(stderr, "fatal: ");
    va_list args;
    va_start(args, fmt);
    vfprintf(stderr, fmt, args);
    va_end(args);
    exit(EXIT_FAILURE);
}
#define pg_fassert(cond)                                                                  \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            fprintf(stderr, "fassert failed '%s' at %s:%d\n", #cond, __FILE__, __LINE__); \
            exit(EXIT_FAILURE);                                                           \
        }                                                                                 \
    } while (0)

#ifndef PG_HOSTSIM
#define CUDA_CHECK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            pg_fatal("CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__, cudaGetErrorString(_e)); \
    } while (0)
#endif

// ================================================================= kernels
// Two launches per step:
//   logic_kernel   one WARP per env. All 32 lanes execute the serial game logic redundantly and in
//                  lockstep (every load/store is warp-uniform), and fan out only inside
//                  pg_scan_down, which turns the reference's O(E) entity-collision loops into E/32
//                  ballots. A warp, not a thread, is the unit so unrelated envs never diverge
//                  against each other and dozens of envs per SM hide each other's load latency.
//   render_kernel  one CTA per env: blit-list build + per-pixel gather + packed RGB store
#ifndef PG_LOGIC_WARPS
#define PG_LOGIC_WARPS 2
#endif
#ifndef PG_LOGIC_MIN_BLOCKS
#define PG_LOGIC_MIN_BLOCKS 24
#endif
#ifndef PG_STEP_CHUNKS
#define PG_STEP_CHUNKS 8
#endif
#ifndef PG_AUX_STREAMS
#define PG_AUX_STREAMS 8
#endif
constexpr int kLogicThreads = 32 * PG_LOGIC_WARPS;  // one warp = one env; few warps per CTA so a finished
constexpr int kLogicEnvsPerBlock = PG_LOGIC_WARPS;  // env frees its slot without waiting on many siblings
constexpr int kRenderThreads = 128;
constexpr int kQuads = RES_W * RES_H / 4;

#ifndef PG_HOSTSIM
// Persistent: the grid is sized to fill the machine once and every warp pulls env indices from a
// global ticket counter until the launch's range is exhausted, so a long env (level reset) only
// delays its own warp and no SM slot idles waiting for a block launch.
template <class G, bool INIT>
__global__ void __launch_bounds__(kLogicThreads, PG_LOGIC_MIN_BLOCKS) logic_kernel(KParams p, unsigned int *ticket) {
    using Frame = typename FrameFor<G>::type;
    const unsigned lane = threadIdx.x & 31u;
    while (true) {
        unsigned t = 0;
        if (lane == 0)
            t = atomicAdd(ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= (unsigned)p.env_count)
            break;
        const int env = p.env_first + (int)t * p.env_step;
        const long long t0 = p.dbg_cycles ? clock64() : 0;
        if (INIT)
            env_init_logic<G, Frame>(p, env);
        else
            env_step_logic<G, Frame>(p, env);
        __syncwarp();
        if (p.dbg_cycles && lane == 0)
            p.dbg_cycles[env] = (uint32_t)(clock64() - t0);
    }
}

#ifndef PG_RENDER_CTAS_PER_SM
#define PG_RENDER_CTAS_PER_SM 0  // 0 = as many as registers / the frame allow
#endif
// Resident CTAs per SM the render kernel is compiled for. The shader is issue-bound and gains from
// occupancy (measured: +24 % on coinrun going from 6 to 8 CTAs/SM = 64 registers), but a frame
// with hundreds of blits does not fit 8 times into shared memory, and there the register cap only
// costs spills.
template <class G>
struct RenderTune {
    static constexpr size_t kFrameBytes = sizeof(typename FrameFor<G>::type);
#ifdef PG_RENDER_MIN_BLOCKS
    static constexpr int kMinBlocks = PG_RENDER_MIN_BLOCKS;
#else
    static constexpr int kMinBlocks = kFrameBytes <= 27 * 1024 ? 8 : (kFrameBytes <= 36 * 1024 ? 6 : 1);
#endif
};

template <class G>
__global__ void __launch_bounds__(kRenderThreads, RenderTune<G>::kMinBlocks) render_kernel(KParams p) {
    using Frame = typename FrameFor<G>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Frame &f = *reinterpret_cast<Frame *>(smem_raw);
    const int env = p.env_first + (int)blockIdx.x * p.env_step;
    const int tid = (int)threadIdx.x;
#ifdef PG_PHASE_TIMING
    long long t0 = clock64(), t1;
#define PG_RENDER_PHASE(id)                                       \
    do {                                                          \
        t1 = clock64();                                           \
        if (tid == 0)                                             \
            p.hdr[env].dbg_phase[id] = (uint32_t)(t1 - t0);       \
        t0 = t1;                                                  \
    } while (0)
#else
#define PG_RENDER_PHASE(id) do { } while (0)
#endif
    env_render_begin<G, Frame>(p, env, f, tid, kRenderThreads);
    __syncthreads();
    PG_RENDER_PHASE(8);
    env_render_build<G, Frame>(p, env, f, tid, kRenderThreads, 32);
    __syncthreads();
    PG_RENDER_PHASE(9);
    if (f.n_jobs > 0) {  // block-uniform
        env_render_tiles<G, Frame>(p, env, f, tid, kRenderThreads);
        __syncthreads();
    }
    if (G::DEFER_ROTATED) {  // compile-time, per game
        env_render_rots<G, Frame>(p, env, f, tid, kRenderThreads);
        __syncthreads();
    }
    env_render_masks<G, Frame>(p, env, f, tid, kRenderThreads);
    __syncthreads();
    PG_RENDER_PHASE(10);
    env_render_pixels<G, Frame>(p, env, f, tid, kRenderThreads);
    PG_RENDER_PHASE(11);
#undef PG_RENDER_PHASE
}
#endif

#ifndef PG_HOSTSIM
// Game::observe without a step (set_state, vecgame.cpp:454-456): camera, then the render kernel
template <class G>
__global__ void camera_kernel(KParams p) {
    using Frame = typename FrameFor<G>::type;
    if (threadIdx.x == 0 && blockIdx.x < (unsigned)p.env_count) {
        const int env = p.env_first + (int)blockIdx.x * p.env_step;
        Ctx c = make_ctx(p, env);
        Raster<G, Frame>::prepare_camera(c);
        write_step_outputs(p, env, *c.h);  // Game::observe's scalar stores, game.cpp:160-164
    }
}
#endif

struct LaunchCtx {
#ifndef PG_HOSTSIM
    cudaStream_t stream;
    cudaStream_t logic_stream;  // null, or a higher-priority stream the logic kernel goes to (then `link` orders render behind it)
    cudaEvent_t link;
    unsigned int *ticket;     // work counter of this launch slot (one per in-flight logic kernel)
    int max_logic_blocks;     // SM count x resident CTAs per SM
    int render_smem_floor;    // dynamic shared memory requested per render CTA is at least this (co-residency knob)
    cudaEvent_t *tev;         // optional: 3 events (before logic, between, after render) for kernel timing
#endif
    int64_t *launch_counter;
};

#ifndef PG_HOSTSIM
// Dynamic shared memory of one render CTA (frame, or the co-residency floor) with the kernel's
// opt-in limit raised to it once.
template <class G>
int prepare_render_smem(const LaunchCtx &lc) {
    using Frame = typename FrameFor<G>::type;
    const int bytes = (int)sizeof(Frame) > lc.render_smem_floor ? (int)sizeof(Frame) : lc.render_smem_floor;
    static int attr_set = 0;
    if (attr_set < bytes) {
        CUDA_CHECK(cudaFuncSetAttribute(render_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        attr_set = bytes;
    }
    return bytes;
}
#endif

template <class G, bool INIT>
void launch_env_kernel(const KParams &p, const LaunchCtx &lc) {
    using Frame = typename FrameFor<G>::type;
    if (p.env_count <= 0)
        return;
#ifndef PG_HOSTSIM
    // Shared memory per render CTA: the frame, or more when the handle asks for fewer resident
    // render CTAs per SM. At 8 CTAs x 128 threads x 64 registers the render kernel owns the whole
    // register file of an SM and no logic-kernel block of another env chunk can run beside it;
    // capping its residency trades a little render speed for real overlap of the two kernels.
    const int render_smem = prepare_render_smem<G>(lc);
    int logic_blocks = (p.env_count + kLogicEnvsPerBlock - 1) / kLogicEnvsPerBlock;
    if (logic_blocks > lc.max_logic_blocks)
        logic_blocks = lc.max_logic_blocks;
    cudaStream_t ls = lc.logic_stream ? lc.logic_stream : lc.stream;
    CUDA_CHECK(cudaMemsetAsync(lc.ticket, 0, sizeof(unsigned int), ls));
    if (lc.tev)
        CUDA_CHECK(cudaEventRecord(lc.tev[0], ls));
    logic_kernel<G, INIT><<<logic_blocks, kLogicThreads, 0, ls>>>(p, lc.ticket);
    if (lc.logic_stream) {
        CUDA_CHECK(cudaEventRecord(lc.link, ls));
        CUDA_CHECK(cudaStreamWaitEvent(lc.stream, lc.link, 0));
    }
    if (lc.tev)
        CUDA_CHECK(cudaEventRecord(lc.tev[1], lc.stream));
    render_kernel<G><<<p.env_count, kRenderThreads, render_smem, lc.stream>>>(p);
    if (lc.tev)
        CUDA_CHECK(cudaEventRecord(lc.tev[2], lc.stream));
    CUDA_CHECK(cudaGetLastError());
    (*lc.launch_counter) += 2;
#else
    static thread_local Frame *f = new Frame;
    for (int b = 0; b < p.env_count; b++) {
        int env = p.env_first + b * p.env_step;
        if (INIT)
            env_init_logic<G, Frame>(p, env);
        else
            env_step_logic<G, Frame>(p, env);
        env_render_begin<G, Frame>(p, env, *f, 0, 1);
        env_render_build<G, Frame>(p, env, *f, 0, 1, 1);
        env_render_masks<G, Frame>(p, env, *f, 0, 1);
        for (int quad = 0; quad < kQuads; quad++) env_render_quad<G, Frame>(p, env, *f, quad);
    }
    (*lc.launch_counter) += 2;
#endif
}

template <class G>
void launch_observe_only(const KParams &p, const LaunchCtx &lc) {
    using Frame = typename FrameFor<G>::type;
    if (p.env_count <= 0)
        return;
#ifndef PG_HOSTSIM
    const int render_smem = prepare_render_smem<G>(lc);
    camera_kernel<G><<<p.env_count, 32, 0, lc.stream>>>(p);
    render_kernel<G><<<p.env_count, kRenderThreads, render_smem, lc.stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
#else
    static thread_local Frame *f = new Frame;
    for (int b = 0; b < p.env_count; b++) {
        int env = p.env_first + b * p.env_step;
        Ctx c = make_ctx(p, env);
        Raster<G, Frame>::prepare_camera(c);
        write_step_outputs(p, env, *c.h);
        env_render_begin<G, Frame>(p, env, *f, 0, 1);
        env_render_build<G, Frame>(p, env, *f, 0, 1, 1);
        env_render_masks<G, Frame>(p, env, *f, 0, 1);
        for (int quad = 0; quad < kQuads; quad++) env_render_quad<G, Frame>(p, env, *f, quad);
    }
#endif
    (*lc.launch_counter) += 2;
}

struct GameVTable {
    const char *name;
    int id;
    int ent_cap, grid_cap, scratch_words;
    int rot_records;  // rotated-sprite records kept in global memory per env (0 = the frame holds them)
    void (*init)(const KParams &, const LaunchCtx &);
    void (*step)(const KParams &, const LaunchCtx &);
    void (*observe_only)(const KParams &, const LaunchCtx &);
};

template <class G>
GameVTable make_vtable(int id) {
    return GameVTable{G::NAME, id, G::ENT_CAP, G::GRID_CAP, G::SCRATCH_WORDS, FrameFor<G>::type::kRotInGlobal ? G::MAX_ROT_BLITS : 0,
                      &launch_env_kernel<G, true>, &launch_env_kernel<G, false>, &launch_observe_only<G>};
}


}  // namespace pg
