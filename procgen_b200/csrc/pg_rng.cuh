// MT19937 as a plain struct in HBM, bit-compatible with libstdc++'s std::mt19937, which is the
// reference's only randomness source (randgen.h:14, randgen.cpp:6-93).
//
// B200 layout choice: instead of regenerating all 624 words every 624 draws (a 624-iteration
// serial burst in whichever step crosses the boundary), each word is twisted on demand the first
// time it is consumed in a generation.  `gen` records how far the current generation has been
// regenerated, so a state imported from a libstdc++ text dump (fully regenerated, gen = 624)
// and a native one produce the same stream.  One draw touches 3 words + 1 store.
#pragma once
#include "pg_common.cuh"

namespace pg {

struct MT19937 {
    uint32_t mt[624];
    int32_t p;        // next word to hand out (624 = generation exhausted), == libstdc++ _M_p
    int32_t gen;      // words [0, gen) of the current generation are already twisted
    int32_t seeded;   // RandGen::is_seeded
    int32_t pad;
};

PG_HD void mt_seed(MT19937 &s, uint32_t seed) {
    // std::mersenne_twister_engine::seed(value): x[i] = f * (x[i-1] ^ (x[i-1] >> 30)) + i
    uint32_t prev = seed;
    s.mt[0] = prev;
    for (int i = 1; i < 624; i++) {
        prev = 1812433253u * (prev ^ (prev >> 30)) + (uint32_t)i;
        s.mt[i] = prev;
    }
    s.p = 624;
    s.gen = 624;
    s.seeded = 1;
}

PG_HD uint32_t mt_next(MT19937 &s) {
    if (s.p >= 624) {
        s.p = 0;
        s.gen = 0;
    }
    const int k = s.p;
    if (k >= s.gen) {
        const int k1 = (k + 1 == 624) ? 0 : k + 1;
        const int km = (k + 397 >= 624) ? k + 397 - 624 : k + 397;
        const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[k1] & 0x7fffffffu);
        s.mt[k] = s.mt[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        s.gen = k + 1;
    }
    uint32_t z = s.mt[k];
    s.p = k + 1;
    z ^= (z >> 11);
    z ^= (z << 7) & 0x9d2c5680u;
    z ^= (z << 15) & 0xefc60000u;
    z ^= (z >> 18);
    return z;
}

// `count` consecutive raw draws, identical to `count` mt_next() calls. On the device the warp's
// lanes regenerate and temper up to 32 consecutive state words at a time: word k of a generation
// depends on the OLD words k, k+1 and on word k+397 (old for k < 227, already NEW otherwise, and
// then more than 32 positions back), so a batch only has to read everything before it writes.
PG_HD void rand_fill_raw(MT19937 &s, uint32_t *out, int count) {
#if defined(__CUDA_ARCH__)
    const int lane = (int)(threadIdx.x & 31u);
    int done = 0;
    while (done < count) {
        int p0 = s.p, gen0 = s.gen;
        __syncwarp();
        if (p0 >= 624) {
            p0 = 0;
            gen0 = 0;
        }
        int m = 624 - p0;
        if (m > 32)
            m = 32;
        if (m > count - done)
            m = count - done;
        const int k = p0 + lane;
        const bool act = lane < m;
        uint32_t v = 0;
        bool twist = false;
        if (act) {
            twist = k >= gen0;
            if (twist) {
                const int k1 = (k + 1 == 624) ? 0 : k + 1;
                const int km = (k + 397 >= 624) ? k + 397 - 624 : k + 397;
                const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[k1] & 0x7fffffffu);
                v = s.mt[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            } else {
                v = s.mt[k];
            }
        }
        __syncwarp();
        if (act && twist)
            s.mt[k] = v;
        if (act) {
            uint32_t z = v;
            z ^= (z >> 11);
            z ^= (z << 7) & 0x9d2c5680u;
            z ^= (z << 15) & 0xefc60000u;
            z ^= (z >> 18);
            out[done + lane] = z;
        }
        s.p = p0 + m;
        s.gen = (p0 + m > gen0) ? p0 + m : gen0;
        done += m;
        __syncwarp();
    }
#else
    for (int i = 0; i < count; i++) out[i] = mt_next(s);
#endif
}

// randgen.cpp:6-11
PG_HD int rand_randint(MT19937 &s, int low, int high) {
    uint32_t x = mt_next(s);
    uint32_t range = (uint32_t)(high - low);
    return (int)((uint32_t)low + (x % range));
}
// randgen.cpp:13-17
PG_HD int rand_randn(MT19937 &s, int high) {
    uint32_t x = mt_next(s);
    return (int)(x % (uint32_t)high);
}
// randgen.cpp:19-23
PG_HD float rand_rand01(MT19937 &s) {
    uint32_t x = mt_next(s);
    return (float)((double)(x) / 4294967296.0);
}
// randgen.cpp:25-27
PG_HD bool rand_randbool(MT19937 &s) { return rand_rand01(s) > .5; }
// randgen.cpp:29-31
PG_HD float rand_randrange(MT19937 &s, float low, float high) { return rand_rand01(s) * (high - low) + low; }
// randgen.cpp:95-98 (randint() with no arguments returns the raw draw as int)
PG_HD int rand_randint_raw(MT19937 &s) { return (int)mt_next(s); }

}  // namespace pg
