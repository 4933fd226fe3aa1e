// Checks pg_atan2f (procgen_b200/csrc/pg_common.cuh) bit for bit against the host C library's
// atan2f — the function the oracle's Entity::face_direction resolves to. Built and run by
// tests/test_device_code_on_cpu.py; prints "OK <cases>" or the first mismatches.
#include <cmath>
#include <cstdio>
#include <cstring>
#include "pg_common.cuh"

static int32_t bits(float f) {
    int32_t i;
    memcpy(&i, &f, 4);
    return i;
}

int main() {
    long bad = 0, n = 0;
    unsigned s = 12345;
    for (long i = 0; i < 6000000; i++) {
        s = s * 1664525u + 1013904223u;
        unsigned a = s;
        s = s * 1664525u + 1013904223u;
        unsigned b = s;
        float x, y;
        if (i & 1) {  // game-scale velocities and offsets
            x = ((int)(a >> 8) - (1 << 23)) * (1.0f / (1 << 20));
            y = ((int)(b >> 8) - (1 << 23)) * (1.0f / (1 << 21));
        } else {  // arbitrary bit patterns
            memcpy(&x, &a, 4);
            memcpy(&y, &b, 4);
            if (x != x || y != y)
                continue;
        }
        float p = atan2f(y, x), q = pg::pg_atan2f(y, x);
        n++;
        if (bits(p) != bits(q)) {
            if (bad < 10)
                printf("MISMATCH y=%a x=%a libm %a mine %a\n", y, x, p, q);
            bad++;
        }
    }
    const float sp[] = {0.0f, -0.0f, 1.0f, -1.0f, 0.5f, -0.5f, 2.0f, 1e30f, -1e30f, 1e-30f, INFINITY, -INFINITY, 3.0f, -0.05f, 0.05f, 0.8f, -0.8f};
    const int ns = sizeof(sp) / sizeof(sp[0]);
    for (int i = 0; i < ns; i++)
        for (int j = 0; j < ns; j++) {
            float p = atan2f(sp[i], sp[j]), q = pg::pg_atan2f(sp[i], sp[j]);
            n++;
            if (bits(p) != bits(q)) {
                printf("MISMATCH special y=%a x=%a libm %a mine %a\n", sp[i], sp[j], p, q);
                bad++;
            }
        }
    if (bad) {
        printf("FAILED %ld of %ld\n", bad, n);
        return 1;
    }
    printf("OK %ld\n", n);
    return 0;
}
