// Cellular-automaton cave generator on per-env scratch, restating roomgen.cpp.
// std::set<int> -> membership flags (ascending scan = sorted iteration), std::queue -> ring in an
// array. The automaton update is per-cell independent and runs across the warp's lanes; the
// breadth-first searches are order-sensitive (they feed RNG indices) and stay serial.
#pragma once
#include "pg_engine.cuh"

namespace pg {

template <class G>
struct RoomGen {
    using E = Engine<G>;
    Ctx *c;
    int n;               // grid_size
    int32_t *next_cells; // [n]
    int32_t *all_rooms;  // [n] flags
    int32_t *room;       // [n] flags (next_room)
    int32_t *queue;      // [4n]
    int32_t *parents;    // [4n]
    int32_t *covered;    // [n] flags
    bool ok;

    static PG_HD int words_needed(int grid_size) { return grid_size * 12; }

    PG_HD void init(Ctx &ctx, int32_t *base, int cap_words) {
        c = &ctx;
        n = ctx.h->grid_size;
        ok = words_needed(n) <= cap_words;
        if (!ok) {
            ctx.h->err |= ERR_SCRATCH_OVERFLOW;
            n = 0;
        }
        next_cells = base;
        all_rooms = base + n;
        room = base + 2 * n;
        covered = base + 3 * n;
        queue = base + 4 * n;
        parents = base + 8 * n;
    }

    // roomgen.cpp:3-19
    static PG_HD int count_neighbors(Ctx &ctx, int idx, int type) {
        const int w = ctx.mw;
        const int x = idx % w, y = idx / w;
        int neighbors = 0;
        for (int i = -1; i <= 1; i++)
            for (int j = -1; j <= 1; j++)
                if (E::get_obj(ctx, x + i, y + j) == type)
                    neighbors++;
        return neighbors;
    }
    // roomgen.cpp:21-36
    PG_HD void update() {
        Ctx *cp = c;
        int32_t *nc = next_cells;
        pg_warp_for(n, [=](int i) { nc[i] = count_neighbors(*cp, i, WALL_OBJ) >= 5 ? WALL_OBJ : SPACE; });
        int16_t *g = c->grid;
        pg_warp_for(n, [=](int i) { g[i] = (int16_t)nc[i]; });
    }
    // roomgen.cpp:38-69: flood from idx; members get `stamp` in `room`, returns the member count.
    // Rooms are connected components, hence disjoint: a cell carrying an older stamp can never be
    // reached from a seed outside its room, so "not yet in THIS room" is room[cell] != stamp and the
    // array needs no clearing between rooms.
    PG_HD int build_room(int idx, int stamp) {
        Ctx &ctx = *c;
        if (E::get_obj_idx(ctx, idx) != SPACE)
            return 0;
        int head = 0, tail = 0, count = 0;
        queue[tail++] = idx;
        const int w = ctx.mw;
        while (head < tail) {
            int curr_idx = queue[head++];
            if (E::get_obj_idx(ctx, curr_idx) != SPACE)
                continue;
            int x = curr_idx % w, y = curr_idx / w;
            for (int i = -1; i <= 1; i++) {
                for (int j = -1; j <= 1; j++) {
                    if ((i == 0 || j == 0) && (i + j != 0)) {
                        int next_idx = E::to_grid_idx(ctx, x + i, y + j);
                        if (next_idx < 0)
                            continue;  // INVALID_IDX: get_obj gives the out-of-bounds object, never SPACE here
                        if (room[next_idx] != stamp && E::get_obj_idx(ctx, next_idx) == SPACE) {
                            if (tail < 4 * n)
                                queue[tail++] = next_idx;
                            room[next_idx] = stamp;
                            count++;
                        }
                    }
                }
            }
        }
        return count;
    }
    // roomgen.cpp:126-145: result flags in `best` (caller buffer [n]); returns its size. The
    // reference builds a std::set per room and unions them into all_rooms; here every room stamps
    // its cells with its ordinal, "already in some room" is a non-zero stamp, and the winner (first
    // room of the largest size, as in the reference's strict `>`) is materialised once at the end.
    PG_HD int find_best_room(int32_t *best) {
        Ctx &ctx = *c;
        int32_t *rm = room;
        pg_warp_for(n, [=](int i) { rm[i] = 0; });
        int best_room_size = -1;
        int best_stamp = 0;
        int stamp = 0;
        for (int i = 0; i < n; i++) {
            if (E::get_obj_idx(ctx, i) == SPACE && !room[i]) {
                stamp++;
                int sz = build_room(i, stamp);
                if (sz > best_room_size) {
                    best_room_size = sz;
                    best_stamp = stamp;
                }
            }
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        pg_warp_for(n, [=](int k) { best[k] = (best_stamp != 0 && rm[k] == best_stamp) ? 1 : 0; });
        return best_room_size < 0 ? 0 : best_room_size;
    }
    // roomgen.cpp:71-124: BFS path src -> dst written to `path` (caller buffer), returns its length
    PG_HD int find_path(int src, int dst, int32_t *path) {
        Ctx &ctx = *c;
        if (E::get_obj_idx(ctx, src) != SPACE)
            return 0;
        int32_t *cv = covered;
        pg_warp_for(n, [=](int i) { cv[i] = 0; });
        int size = 0;
        queue[size] = src;
        parents[size] = -1;
        size++;
        int search_idx = 0;
        const int w = ctx.mw;
        while (search_idx < size) {
            int curr_idx = queue[search_idx];
            if (curr_idx == dst)
                break;
            int x = curr_idx % w, y = curr_idx / w;
            for (int i = -1; i <= 1; i++) {
                for (int j = -1; j <= 1; j++) {
                    if ((i == 0 || j == 0) && (i + j != 0)) {
                        int next_idx = E::to_grid_idx(ctx, x + i, y + j);
                        if (next_idx < 0)
                            continue;
                        if (!covered[next_idx] && E::get_obj_idx(ctx, next_idx) == SPACE) {
                            if (size < 4 * n) {
                                queue[size] = next_idx;
                                parents[size] = search_idx;
                                size++;
                            }
                            covered[next_idx] = 1;
                        }
                    }
                }
            }
            search_idx++;
        }
        int len = 0;
        if (search_idx < size && queue[search_idx] == dst) {
            // walk the parent chain, then reverse
            int k = search_idx;
            while (k >= 0) {
                path[len++] = queue[k];
                k = parents[k];
            }
            for (int a = 0, b = len - 1; a < b; a++, b--) {
                int t = path[a];
                path[a] = path[b];
                path[b] = t;
            }
        }
        return len;
    }
    // roomgen.cpp:147-177: grow `set` (flags [n]) by n_iter rings of SPACE cells (8-neighbourhood)
    PG_HD void expand_room(int32_t *set, int n_iter) {
        Ctx &ctx = *c;
        int32_t *curr = room, *next = all_rooms;  // reuse as scratch sets
        pg_warp_for(n, [=](int i) { curr[i] = set[i]; });
        const int w = ctx.mw;
        for (int loop = 0; loop < n_iter; loop++) {
            pg_warp_for(n, [=](int i) { next[i] = 0; });
            for (int curr_idx = 0; curr_idx < n; curr_idx++) {
                if (!curr[curr_idx])
                    continue;
                if (E::get_obj_idx(ctx, curr_idx) != SPACE)
                    continue;
                int x = curr_idx % w, y = curr_idx / w;
                for (int i = -1; i <= 1; i++) {
                    for (int j = -1; j <= 1; j++) {
                        if (i != 0 || j != 0) {
                            int next_idx = E::to_grid_idx(ctx, x + i, y + j);
                            if (next_idx < 0)
                                continue;
                            if (!set[next_idx] && E::get_obj_idx(ctx, next_idx) == SPACE) {
                                set[next_idx] = 1;
                                next[next_idx] = 1;
                            }
                        }
                    }
                }
            }
            int32_t *t = curr;
            curr = next;
            next = t;
        }
    }
};

}  // namespace pg
