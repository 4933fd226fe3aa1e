// Checks procgen_b200/csrc/pg_stdsort.cuh against the C++ library's std::sort, including the order
// of ties and the heapsort fallback (reached with McIlroy's adversarial comparator input).
// Built and run by tests/test_device_code_on_cpu.py; prints "OK <cases>" or the first mismatch.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include "pg_stdsort.cuh"

struct Rec {
    int key, id;
};

static bool check(const std::vector<int> &keys, const char *what) {
    const int n = (int)keys.size();
    std::vector<Rec> recs(n);
    for (int i = 0; i < n; i++) recs[i] = {keys[i], i};
    std::vector<const Rec *> lib(n);
    std::vector<int32_t> mine(n);
    for (int i = 0; i < n; i++) {
        lib[i] = &recs[i];
        mine[i] = i;
    }
    std::sort(lib.begin(), lib.end(), [](const Rec *x, const Rec *y) { return x->key > y->key; });
    pg::pg_std_sort(mine.data(), n, [&](int32_t x, int32_t y) { return recs[x].key > recs[y].key; });
    for (int i = 0; i < n; i++)
        if (lib[i]->id != mine[i]) {
            printf("MISMATCH %s n=%d at %d: lib id %d mine %d\n", what, n, i, lib[i]->id, mine[i]);
            return false;
        }
    return true;
}

// M. D. McIlroy, "A Killer Adversary for Quicksort" (1999): values are decided while sorting
static std::vector<int> g_val;
static int g_nsolid, g_candidate, g_gas;
static bool adversary_less(int x, int y) {
    if (g_val[x] == g_gas && g_val[y] == g_gas) {
        if (x == g_candidate)
            g_val[x] = g_nsolid++;
        else
            g_val[y] = g_nsolid++;
    }
    if (g_val[x] == g_gas)
        g_candidate = x;
    else if (g_val[y] == g_gas)
        g_candidate = y;
    return g_val[x] < g_val[y];
}
static std::vector<int> killer(int n) {
    g_val.assign(n, n);
    g_gas = n;
    g_nsolid = 0;
    g_candidate = 0;
    std::vector<int> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), adversary_less);
    return g_val;
}

int main() {
    std::mt19937 gen(1234);
    int cases = 0;
    for (int n = 0; n <= 300; n++) {
        for (int rep = 0; rep < 20; rep++) {
            std::vector<int> keys(n);
            int spread = rep < 5 ? 3 : (rep < 10 ? 17 : (rep < 15 ? n + 1 : 1000000));
            for (auto &k : keys) k = (int)(gen() % (unsigned)spread);
            if (rep % 4 == 1) std::sort(keys.begin(), keys.end());
            if (rep % 4 == 2) std::sort(keys.begin(), keys.end(), [](int a, int b) { return a > b; });
            if (!check(keys, "random")) return 1;
            cases++;
        }
    }
    for (int n : {17, 33, 64, 100, 245, 256, 1000, 5000}) {
        std::vector<int> k = killer(n);
        // the comparator under test is '>' so feed the mirrored values too
        std::vector<int> neg(k);
        for (auto &v : neg) v = -v;
        std::vector<int> tied(k);
        for (auto &v : tied) v = -(v / 3);
        if (!check(k, "killer") || !check(neg, "killer-mirrored") || !check(tied, "killer-tied")) return 1;
        cases += 3;
    }
    printf("OK %d\n", cases);
    return 0;
}
