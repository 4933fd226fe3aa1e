// Kruskal maze generator on per-env scratch memory (int32 words), restating mazegen.cpp.
//
// The reference keeps std::set<int> per cell and merges them; only set *identity* is observable
// (cell_sets_idxs lookups), so sets are represented by a label per cell and a merge is a relabel,
// done by the whole warp. std::vector::erase on the wall list and the sorted iteration of
// std::set<int> in expand_to_type are observable (they feed RNG indices) and are reproduced
// exactly: order-preserving erase, ascending-index iteration over membership flags.
#pragma once
#include "pg_engine.cuh"

namespace pg {

constexpr int MAZE_OFFSET = 1;

struct MazeGen {
    MT19937 *rng;
    int maze_dim, array_dim, num_free_cells;
    int32_t *grid;        // [array_dim * array_dim], Grid<int> (x, y) -> y * array_dim + x
    int32_t *set_idx;     // [array_dim * array_dim] cell_sets_idxs (indexed maze_dim*y + x)
    int32_t *free_cells;  // [array_dim * array_dim]
    int32_t *free_flag;   // [array_dim * array_dim] free_cell_set membership
    int32_t *walls;       // [4 * max_walls]
    int32_t *set_a;       // [array_dim * array_dim] membership flags for s0 / curr / next / s1
    int32_t *set_b;
    int32_t *set_c;
    int32_t *set_d;
    int32_t *list;        // [array_dim * array_dim] scratch vector
    bool ok;

    static PG_HD int words_needed(int maze_dim) {
        int a = (maze_dim + 2) * (maze_dim + 2);
        return a * 9 + 4 * (maze_dim * maze_dim / 2 + 8);
    }

    // mazegen.cpp:12-20
    PG_HD void init(Ctx &c, int _maze_dim) {
        rng = c.rng;
        maze_dim = _maze_dim;
        array_dim = maze_dim + 2;
        num_free_cells = 0;
        const int a = array_dim * array_dim;
        ok = words_needed(maze_dim) <= c.scratch_cap;
        if (!ok) {
            c.h->err |= ERR_SCRATCH_OVERFLOW;
            maze_dim = 1;
            array_dim = 3;
        }
        int32_t *p = c.scratch;
        grid = p; p += a;
        set_idx = p; p += a;
        free_cells = p; p += a;
        free_flag = p; p += a;
        set_a = p; p += a;
        set_b = p; p += a;
        set_c = p; p += a;
        set_d = p; p += a;
        list = p; p += a;
        walls = p;
        int32_t *g = grid, *fc = free_cells;
        pg_warp_for(a, [=](int i) {
            g[i] = 0;
            fc[i] = 0;
        });
    }

    PG_HD int grid_get(int x, int y) const { return grid[y * array_dim + x]; }
    PG_HD void grid_set(int x, int y, int v) { grid[y * array_dim + x] = v; }
    PG_HD int lookup(int x, int y) const { return set_idx[maze_dim * y + x]; }

    // mazegen.cpp:26-34
    PG_HD void set_free_cell(int x, int y) {
        grid_set(x + MAZE_OFFSET, y + MAZE_OFFSET, SPACE);
        int cell = maze_dim * y + x;
        if (!free_flag[cell]) {
            free_cells[num_free_cells] = cell;
            free_flag[cell] = 1;
            num_free_cells += 1;
        }
    }
    // mazegen.cpp:36-46
    PG_HD int get_obj(int idx) const {
        int x = idx % array_dim;
        int y = idx / array_dim;
        if (x <= 0 || x >= array_dim - 1)
            return INVALID_OBJ;
        if (y <= 0 || y >= array_dim - 1)
            return INVALID_OBJ;
        return grid[y * array_dim + x];
    }
    // mazegen.cpp:48-67 — neighbour order: (-1,0), (0,-1), (0,+1), (+1,0)
    PG_HD int get_neighbors(int idx, int type, int *out) const {
        int x = idx % array_dim;
        int y = idx / array_dim;
        int n = 0;
        for (int dx = -1; dx <= 1; dx++) {
            for (int dy = -1; dy <= 1; dy++) {
                if (dx == 0 && dy == 0)
                    continue;
                if (dx != 0 && dy != 0)
                    continue;
                int n_idx = (y + dy) * array_dim + (x + dx);
                if (get_obj(n_idx) == type)
                    out[n++] = n_idx;
            }
        }
        return n;
    }

    // mazegen.cpp:112-187
    PG_HD void generate_maze() {
        const int a = array_dim * array_dim;
        {
            int32_t *g = grid, *ff = free_flag, *si = set_idx;
            pg_warp_for(a, [=](int i) {
                g[i] = WALL_OBJ;
                ff[i] = 0;
                si[i] = i;
            });
        }
        grid_set(MAZE_OFFSET, MAZE_OFFSET, 0);
        num_free_cells = 0;
        int nwalls = 0;
        for (int i = 1; i < maze_dim; i += 2)
            for (int j = 0; j < maze_dim; j += 2)
                if (i > 0 && i < maze_dim - 1) {
                    int32_t *w = walls + 4 * nwalls++;
                    w[0] = i - 1; w[1] = j; w[2] = i + 1; w[3] = j;
                }
        for (int i = 0; i < maze_dim; i += 2)
            for (int j = 1; j < maze_dim; j += 2)
                if (j > 0 && j < maze_dim - 1) {
                    int32_t *w = walls + 4 * nwalls++;
                    w[0] = i; w[1] = j - 1; w[2] = i; w[3] = j + 1;
                }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        while (nwalls > 0) {
            int n = rand_randn(*rng, nwalls);
            const int32_t *w = walls + 4 * n;
            const int x1 = w[0], y1 = w[1], x2 = w[2], y2 = w[3];
            int s0_idx = lookup(x1, y1);
            int s1_idx = lookup(x2, y2);
            int x0 = (x1 + x2) / 2;
            int y0 = (y1 + y2) / 2;
            int center = maze_dim * y0 + x0;
            bool can_remove = (grid_get(x0 + MAZE_OFFSET, y0 + MAZE_OFFSET) == WALL_OBJ) && (s0_idx != s1_idx);
            if (can_remove) {
                set_free_cell(x1, y1);
                set_free_cell(x0, y0);
                set_free_cell(x2, y2);
#if defined(__CUDA_ARCH__)
                __syncwarp();
#endif
                // s1 <- s1 U s0 U {center}; every member now maps to s1_idx
                int32_t *si = set_idx;
                pg_warp_for(maze_dim * maze_dim, [=](int k) {
                    if (si[k] == s0_idx || k == center)
                        si[k] = s1_idx;
                });
            }
            pg_warp_erase(walls, n, nwalls, 4);
            nwalls--;
        }
    }

    // mazegen.cpp:190-210
    PG_HD void generate_maze_no_dead_ends() {
        generate_maze();
        int adj_space[4], adj_wall[4];
        for (int i = 0; i < array_dim * array_dim; i++) {
            if (get_obj(i) == SPACE) {
                int ns = get_neighbors(i, SPACE, adj_space);
                if (ns == 1) {
                    int nw = get_neighbors(i, WALL_OBJ, adj_wall);
                    if (nw > 0) {
                        int n = rand_randn(*rng, nw);
                        grid[adj_wall[n]] = SPACE;
                    }
                }
            }
        }
    }

    // ---- sets of cell indices as membership flags; iteration is ascending index = std::set<int> order
    PG_HD void set_clear(int32_t *s) {
        pg_warp_for(array_dim * array_dim, [=](int i) { s[i] = 0; });
    }
    PG_HD void set_copy(int32_t *dst, const int32_t *src) {
        pg_warp_for(array_dim * array_dim, [=](int i) { dst[i] = src[i]; });
    }
    PG_HD void set_union(int32_t *dst, const int32_t *src) {
        pg_warp_for(array_dim * array_dim, [=](int i) { dst[i] |= src[i]; });
    }

    // mazegen.cpp:69-98. s0/s1 are membership arrays; returns the first neighbour of `type` found
    // while flooding SPACE outward from s0 (new cells are recorded in s1), or -1.
    PG_HD int expand_to_type(int32_t *s0, int32_t *s1, int type) {
        const int a = array_dim * array_dim;
        int32_t *curr = set_c, *next = set_d;
        set_copy(curr, s0);
        int target_elems[4], adj_space[4];
        while (true) {
            int curr_size = 0;
            set_clear(next);
            for (int elem = 0; elem < a; elem++) {
                if (!curr[elem])
                    continue;
                curr_size++;
                int nt = get_neighbors(elem, type, target_elems);
                int ns = get_neighbors(elem, SPACE, adj_space);
                for (int k = 0; k < ns; k++) {
                    int j = adj_space[k];
                    if (!s0[j] && !s1[j]) {
                        next[j] = 1;
                        s1[j] = 1;
                    }
                }
                if (nt > 0)
                    return target_elems[0];
            }
            if (curr_size == 0)
                break;
            int32_t *t = curr;
            curr = next;
            next = t;
            // (the reference's loop test is curr.size() > 0 on the NEW curr; an empty new curr ends
            // the loop on the next pass through the counter above)
        }
        return -1;
    }

    // randgen.cpp:53-70 on a list held in `list` (n entries): returns chosen count, results in out[]
    PG_HD int choose_n(int n, int k, int *out) {
        if (k > n) {
            for (int i = 0; i < n; i++) out[i] = list[i];
            return n;
        }
        int rem = n;
        int chosen = 0;
        while (chosen < k) {
            int idx = rand_randn(*rng, rem);
            out[chosen++] = list[idx];
            for (int j = idx; j < rem - 1; j++) list[j] = list[j + 1];
            rem--;
        }
        return chosen;
    }

    // mazegen.cpp:213-290
    PG_HD void generate_maze_with_doors(int num_doors) {
        generate_maze();
        const int a = array_dim * array_dim;
        int adj[4];
        int nforks = 0;
        for (int i = 0; i < a; i++) {
            if (get_obj(i) == SPACE) {
                int ns = get_neighbors(i, SPACE, adj);
                if (ns > 2)
                    list[nforks++] = i;
            }
        }
        int chosen[8];
        if (num_doors > 8)
            num_doors = 8;
        num_doors = choose_n(nforks, num_doors, chosen);
        for (int i = 0; i < num_doors; i++) grid[chosen[i]] = DOOR_OBJ;

        int agent_cell;
        {
            int nspace = 0;
            for (int i = 0; i < a; i++)
                if (get_obj(i) == SPACE)
                    list[nspace++] = i;
            int nd;
            do {
                agent_cell = list[rand_randn(*rng, nspace)];
                nd = get_neighbors(agent_cell, DOOR_OBJ, adj);
            } while (nd > 0);
            grid[agent_cell] = AGENT_OBJ;
        }

        int32_t *s0 = set_a, *s1 = set_b;
        set_clear(s0);
        s0[agent_cell] = 1;
        for (int door_num = 0; door_num < num_doors + 1; door_num++) {
            set_clear(s1);
            int found_door = -1;
            if (door_num < num_doors) {
                found_door = expand_to_type(s0, s1, DOOR_OBJ);
                if (found_door >= 0)
                    grid[found_door] = DOOR_OBJ + door_num + 1;
                set_union(s0, s1);
            }
            expand_to_type(s0, s1, -999);
            int nspace = 0;
            for (int x = 0; x < a; x++)
                if (s1[x])
                    list[nspace++] = x;
            if (nspace == 0)
                return;  // reference: fassert(space_cells.size() > 0)
            int key_cell = list[rand_randn(*rng, nspace)];
            grid[key_cell] = door_num == num_doors ? EXIT_OBJ : (KEY_OBJ + door_num + 1);
            set_union(s0, s1);
            if (found_door >= 0)
                s0[found_door] = 1;
        }
    }

    // mazegen.cpp:292-306
    PG_HD void place_objects(int start_obj, int num_objs) {
        for (int j = 0; j < num_objs; j++) {
            int m = rand_randn(*rng, num_free_cells);
            while (free_cells[m] == -1 || free_cells[m] == 0) m = rand_randn(*rng, num_free_cells);
            int coin_cell = free_cells[m];
            free_cells[m] = -1;
            grid_set(coin_cell % maze_dim + MAZE_OFFSET, coin_cell / maze_dim + MAZE_OFFSET, start_obj + j);
        }
    }
};

}  // namespace pg
