// procgen_b200 — shared definitions for the device engine.
//
// Everything marked PG_HD compiles for sm_100a (product) and, for the CPU debugging harness under
// tests/hostsim only, as plain host C++ (g++ -x c++).  The host build exists so the bit-exact
// float/integer semantics can be diffed against the oracle without a GPU in the loop; it is never
// linked into the product library and the product has no CPU path.
//
// Float semantics (SURVEY §7 hard part 1): the reference is C++ that mixes float and double
// freely and whose published wheels are built without FMA (CMakeLists.txt:30).  Device code is
// compiled with -fmad=false and written with the same operand types as the reference expression
// it mirrors; libm calls that the reference resolves to the C `double` overloads (unqualified
// sqrt/floor/ceil/fabs on floats — checked in the oracle's disassembly) go through the pg_d*
// helpers below so nvcc cannot pick the float overloads.
#pragma once

#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#define PG_HD_NOINLINE __host__ __device__ __noinline__
// free (non-member, non-template) functions defined in headers: `inline` for linkage, never inlined on the device
#define PG_HD_FREE_NOINLINE inline __host__ __device__ __noinline__
#define PG_D __device__ __forceinline__
#else
#define PG_HD inline
#define PG_HD_NOINLINE
#define PG_HD_FREE_NOINLINE inline
#define PG_D inline
#endif

namespace pg {

// ---- observation contract (game.h:24-27): constants forever
constexpr int RES_W = 64;
constexpr int RES_H = 64;

// ---- object ids (object-ids.h:9-26)
constexpr int INVALID_OBJ = -1;
constexpr int INVALID_IDX = -2;
constexpr int PLAYER = 0;
constexpr int SPACE = 100;
constexpr int WALL_OBJ = 51;
constexpr int EXIT_OBJ = 52;
constexpr int AGENT_OBJ = 53;
constexpr int EXPLOSION = 54;
constexpr int EXPLOSION2 = 55;
constexpr int EXPLOSION3 = 56;
constexpr int EXPLOSION4 = 57;
constexpr int EXPLOSION5 = 58;
constexpr int TRAIL = 59;
constexpr int DOOR_OBJ = 200;
constexpr int KEY_OBJ = 300;

// ---- engine constants (basic-abstract-game.cpp:6-20, cpp-utils.h:12)
constexpr float PI_F = 3.14159265358979323846264338327950288f;
constexpr float MAXVTHETA = 15 * PI_F / 180;
constexpr float MIXRATEROT = 0.5f;
constexpr float POS_EPS = -0.001f;
constexpr float RENDER_EPS = 0.02f;
constexpr int USE_ASSET_THRESHOLD = 100;
constexpr int MAX_ASSETS = USE_ASSET_THRESHOLD;
constexpr int MAX_IMAGE_THEMES = 10;

// ---- distribution modes (game.h:32-37)
constexpr int EasyMode = 0;
constexpr int HardMode = 1;
constexpr int ExtremeMode = 2;
constexpr int MemoryMode = 10;

// ---- game ids (order = alphabetical, the registry is a name->factory map, game-registry.h)
enum GameId : int {
    GAME_BIGFISH = 0,
    GAME_BOSSFIGHT,
    GAME_CAVEFLYER,
    GAME_CHASER,
    GAME_CLIMBER,
    GAME_COINRUN,
    GAME_DODGEBALL,
    GAME_FRUITBOT,
    GAME_HEIST,
    GAME_JUMPER,
    GAME_LEAPER,
    GAME_MAZE,
    GAME_MINER,
    GAME_NINJA,
    GAME_PLUNDER,
    GAME_STARPILOT,
    NUM_GAMES
};

// ---- double-precision libm shims (the reference calls the C double overloads)
PG_HD double pg_dsqrt(double x) { return sqrt(x); }
PG_HD double pg_dfloor(double x) { return floor(x); }
PG_HD double pg_dceil(double x) { return ceil(x); }
PG_HD double pg_dfabs(double x) { return fabs(x); }

// ---- atan2f exactly as the oracle's C library computes it.
// Entity::face_direction (entity.cpp:84-88) calls atan2 on floats with <math.h> in scope, which
// binds to atan2f (checked in the oracle's object code). glibc 2.39's atan2f/atanf (the fdlibm
// single-precision kernels, sysdeps/ieee754/flt-32/e_atan2f.c + s_atanf.c) are NOT correctly
// rounded — they differ from round(atan2(double)) on ~16% of inputs — so the float operation
// sequence is restated here: same reduction intervals, same coefficient bits, no FMA (device code
// is built with -fmad=false). tests/native/atan2f_check.cpp diffs it against the host libm.
PG_HD int32_t pg_float_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    union { float f; int32_t i; } u;
    u.f = f;
    return u.i;
#endif
}
PG_HD float pg_bits_float(int32_t i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    union { float f; int32_t i; } u;
    u.i = i;
    return u.f;
#endif
}
PG_HD float pg_atanf(float x) {
    const float hi0 = pg_bits_float(0x3eed6338), hi1 = pg_bits_float(0x3f490fda), hi2 = pg_bits_float(0x3f7b985e), hi3 = pg_bits_float(0x3fc90fda);
    const float lo0 = pg_bits_float(0x31ac3769), lo1 = pg_bits_float(0x33222168), lo2 = pg_bits_float(0x33140fb4), lo3 = pg_bits_float(0x33a22168);
    const float aT0 = pg_bits_float(0x3eaaaaab), aT1 = pg_bits_float((int32_t)0xbe4ccccd), aT2 = pg_bits_float(0x3e124925),
                aT3 = pg_bits_float((int32_t)0xbde38e38), aT4 = pg_bits_float(0x3dba2e6e), aT5 = pg_bits_float((int32_t)0xbd9d8795),
                aT6 = pg_bits_float(0x3d886b35), aT7 = pg_bits_float((int32_t)0xbd6ef16b), aT8 = pg_bits_float(0x3d4bda59),
                aT9 = pg_bits_float((int32_t)0xbd15a221), aT10 = pg_bits_float(0x3c8569d7);
    const int32_t hx = pg_float_bits(x);
    const int32_t ix = hx & 0x7fffffff;
    float hi = 0, lo = 0;
    int id;
    if (ix >= 0x4c000000) {  // |x| >= 2^25
        if (ix > 0x7f800000)
            return x + x;
        return hx > 0 ? hi3 + lo3 : -hi3 - lo3;
    }
    if (ix < 0x3ee00000) {  // |x| < 0.4375
        if (ix < 0x31000000)
            return x;
        id = -1;
    } else {
        x = pg_bits_float(ix);
        if (ix < 0x3f980000) {
            if (ix < 0x3f300000) {
                id = 0; hi = hi0; lo = lo0;
                x = (2.0f * x - 1.0f) / (2.0f + x);
            } else {
                id = 1; hi = hi1; lo = lo1;
                x = (x - 1.0f) / (x + 1.0f);
            }
        } else {
            if (ix < 0x401c0000) {
                id = 2; hi = hi2; lo = lo2;
                x = (x - 1.5f) / (1.0f + 1.5f * x);
            } else {
                id = 3; hi = hi3; lo = lo3;
                x = -1.0f / x;
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0)
        return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return hx < 0 ? -r : r;
}
PG_HD float pg_atan2f(float y, float x) {
    const float tiny = 1.0e-30f;
    const float pi_o_4 = pg_bits_float(0x3f490fdb), pi_o_2 = pg_bits_float(0x3fc90fdb), pi = pg_bits_float(0x40490fdb),
                pi_lo = pg_bits_float((int32_t)0xb3bbbd2e);
    const int32_t hx = pg_float_bits(x), hy = pg_float_bits(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000)
        return x + y;
    if (hx == 0x3f800000)
        return pg_atanf(y);
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);  // 2*sign(x) + sign(y)
    if (iy == 0) {
        if (m < 2)
            return y;
        return m == 2 ? pi + tiny : -pi - tiny;
    }
    if (ix == 0)
        return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000) {
        if (iy == 0x7f800000) {
            if (m == 0) return pi_o_4 + tiny;
            if (m == 1) return -pi_o_4 - tiny;
            if (m == 2) return 3.0f * pi_o_4 + tiny;
            return -3.0f * pi_o_4 - tiny;
        }
        if (m == 0) return 0.0f;
        if (m == 1) return -0.0f;
        if (m == 2) return pi + tiny;
        return -pi - tiny;
    }
    if (iy == 0x7f800000)
        return hy < 0 ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int k = (iy - ix) >> 23;
    float z;
    if (k > 60)
        z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60)
        z = 0.0f;
    else
        z = pg_atanf(pg_bits_float(pg_float_bits(y / x) & 0x7fffffff));
    if (m == 0) return z;
    if (m == 1) return pg_bits_float(pg_float_bits(z) ^ (int32_t)0x80000000);
    if (m == 2) return pi - (z - pi_lo);
    return (z - pi_lo) - pi;
}

// cpp-utils.h:43-45 — returns double on purpose
PG_HD double pg_sign(double x) { return x > 0 ? +1 : (x == 0 ? 0 : -1); }

// cpp-utils.h:47-53
PG_HD float pg_clip_abs(float x, float y) {
    if (x > y)
        return y;
    if (x < -y)
        return -y;
    return x;
}

// Qt 6's qRound(double) (qnumeric.h) — half away from zero, used by every raster rule. (Qt 5
// rounded exact negative ties up; the oracle is pinned against Qt 6.6.3.)
PG_HD int pg_qround(double d) {
    return d >= 0.0 ? int(d + 0.5) : int(d - 0.5);
}

// Qt's BYTE_MUL on a packed 0xAARRGGBB word (qdrawhelper_p.h)
PG_HD uint32_t pg_byte_mul(uint32_t x, uint32_t a) {
    uint32_t t = (x & 0xff00ffu) * a;
    t = (t + ((t >> 8) & 0xff00ffu) + 0x800080u) >> 8;
    t &= 0xff00ffu;
    x = ((x >> 8) & 0xff00ffu) * a;
    x = (x + ((x >> 8) & 0xff00ffu) + 0x800080u);
    x &= 0xff00ff00u;
    return x | t;
}

// ---- warp-cooperative scan (logic kernel: one warp = one env, all 32 lanes run the serial game
// logic redundantly and in lockstep — every load/store is warp-uniform — and split up only inside
// these scans, which replace the reference's O(E) "for each entity, test" loops).
// pg_scan_down(upper, pred): largest i in [0, upper) with pred(i), or -1. pred must be read-only.
// Sequential semantics are preserved by the callers: they handle hit i, then rescan [0, i).
template <class F>
PG_HD int pg_scan_down(int upper, F pred) {
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31u);
    for (int base = upper - 1; base >= 0; base -= 32) {
        const int i = base - lane;
        const bool hit = (i >= 0) && pred(i);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m)
            return base - (__ffs((int)m) - 1);
    }
    return -1;
#else
    for (int i = upper - 1; i >= 0; i--)
        if (pred(i))
            return i;
    return -1;
#endif
}

// Descending scan that keeps the rest of a chunk's ballot when handling a hit changed nothing
// the predicate depends on (the common case: overlapping but non-interacting entities such as
// trails); `restart_below(i)` discards it after a hit that moved something.
struct ScanDownIter {
    int next_base;   // highest index of the next chunk to test
    int cur_base;    // highest index of the chunk `mask` belongs to
    unsigned mask;   // device: remaining hits of the current chunk (bit = lane, lane 0 = cur_base)
    PG_HD explicit ScanDownIter(int upper) : next_base(upper - 1), cur_base(upper - 1), mask(0) {}
    template <class F>
    PG_HD int next(F pred) {
#if defined(__CUDA_ARCH__)
        const int lane = (int)(threadIdx.x & 31u);
        while (true) {
            if (mask) {
                const int b = __ffs((int)mask) - 1;
                mask &= mask - 1;
                return cur_base - b;
            }
            if (next_base < 0)
                return -1;
            cur_base = next_base;
            next_base -= 32;
            const int i = cur_base - lane;
            const bool hit = (i >= 0) && pred(i);
            mask = __ballot_sync(0xffffffffu, hit);
        }
#else
        while (next_base >= 0) {
            const int i = next_base--;
            if (pred(i))
                return i;
        }
        return -1;
#endif
    }
    PG_HD void restart_below(int i) {
        mask = 0;
        next_base = i - 1;
    }
};

// Ascending twin of ScanDownIter: smallest i in [from, n) with pred(i), keeping the rest of the
// chunk's ballot until restart_from() is called.
struct ScanUpIter {
    int next_base, cur_base, n;
    unsigned mask;
    PG_HD ScanUpIter(int from, int n_) : next_base(from), cur_base(from), n(n_), mask(0) {}
    template <class F>
    PG_HD int next(F pred) {
#if defined(__CUDA_ARCH__)
        const int lane = (int)(threadIdx.x & 31u);
        while (true) {
            if (mask) {
                const int b = __ffs((int)mask) - 1;
                mask &= mask - 1;
                return cur_base + b;
            }
            if (next_base >= n)
                return -1;
            cur_base = next_base;
            next_base += 32;
            const int i = cur_base + lane;
            const bool hit = (i < n) && pred(i);
            mask = __ballot_sync(0xffffffffu, hit);
        }
#else
        while (next_base < n) {
            const int i = next_base++;
            if (pred(i))
                return i;
        }
        return -1;
#endif
    }
    PG_HD void restart_from(int i) {
        mask = 0;
        next_base = i;
    }
};

// pg_warp_for(n, f): f(i) for every i in [0, n); iterations must be independent (disjoint writes).
// On the device the warp's lanes stride over i and a __syncwarp() publishes the writes before the
// lanes go back to uniform execution.
template <class F>
PG_HD void pg_warp_for(int n, F f) {
#if defined(__CUDA_ARCH__)
    for (int i = (int)(threadIdx.x & 31u); i < n; i += 32) f(i);
    __syncwarp();
#else
    for (int i = 0; i < n; i++) f(i);
#endif
}

// out[] = every i in [0, n) with pred(i), ascending (the reference's "for i: if (...) v.push_back(i)");
// returns the count. pred must be pure. Device: ballot + popcount ranks per 32-index chunk.
template <class F>
PG_HD int pg_warp_compact(int n, int32_t *out, F pred) {
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31u);
    int count = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane;
        const bool hit = (i < n) && pred(i);
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit)
            out[count + __popc(m & ((1u << lane) - 1u))] = i;
        count += __popc(m);
    }
    __syncwarp();
    return count;
#else
    int count = 0;
    for (int i = 0; i < n; i++)
        if (pred(i))
            out[count++] = i;
    return count;
#endif
}

// a[from .. n-2] = a[from+1 .. n-1] for records of `words` int32 each (vector::erase of one element)
PG_HD void pg_warp_erase(int32_t *a, int from, int n, int words) {
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31u);
    const int total = (n - 1 - from) * words;
    int32_t *dst = a + from * words;
    for (int base = 0; base < total; base += 32) {
        const int k = base + lane;
        int32_t v = 0;
        if (k < total)
            v = dst[k + words];
        __syncwarp();
        if (k < total)
            dst[k] = v;
        __syncwarp();
    }
#else
    for (int k = from * words; k < (n - 1) * words; k++) a[k] = a[k + words];
#endif
}

// error bits latched per env (the reference would fassert/exit; we must not kill the GPU)
enum ErrBits : uint32_t {
    ERR_ENTITY_OVERFLOW = 1u << 0,
    ERR_GRID_OOB = 1u << 1,
    ERR_BLIT_OVERFLOW = 1u << 2,
    ERR_SCRATCH_OVERFLOW = 1u << 3,
    ERR_FASSERT = 1u << 4,
    ERR_UNSUPPORTED = 1u << 5,
    ERR_TILE_ARENA = 1u << 6,   // a frame needed more tiles / general cell blits than its shared-memory arena holds
    ERR_ENT_BLITS = 1u << 7,    // more visible entity blits than the frame keeps (MAX_VISIBLE_ENTS)
    ERR_ROT_BLITS = 1u << 8,    // more rotated sprites than the frame keeps (MAX_ROT_BLITS)
};

}  // namespace pg
