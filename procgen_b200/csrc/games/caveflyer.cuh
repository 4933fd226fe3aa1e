// CaveFlyer on the device engine. Behaviour restated from games/caveflyer.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"
#include "../pg_roomgen.cuh"

namespace pg {

struct CaveFlyerGame : Defaults<CaveFlyerGame>, DrawDefaults<CaveFlyerGame> {
    using E = Engine<CaveFlyerGame>;
    static constexpr int ENT_CAP = 192;
    static constexpr int GRID_CAP = 60 * 60;
    static constexpr int SCRATCH_WORDS = 18 * GRID_CAP;
    static constexpr int MAX_VISIBLE_ENTS = 192;
    static constexpr int MAX_ROT_BLITS = 32;
    static constexpr int MAX_VIEW_CELLS = 20;  // visibility 16 centred
    static constexpr int FULL_VIEW_CELLS = 60;  // center_agent = false: the whole world (basic-abstract-game.cpp:819-838)
    static constexpr const char *NAME = "caveflyer";
    // superset of the types is_blocked (:49-55) and will_reflect (:80) accept
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob || t == CAVEWALL; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    // caveflyer.cpp:9-20
    static constexpr float GOAL_REWARD = 10.0f;
    static constexpr float TARGET_REWARD = 3.0f;
    static constexpr int GOAL = 1, OBSTACLE = 2, TARGET = 3, PLAYER_BULLET = 4, ENEMY = 5, CAVEWALL = 8, EXHAUST = 9;
    static constexpr int MARKER = 1003;

    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->mixrate = 0.9f;
    }
    // caveflyer.cpp:56-70
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        int t = c.ents[oi].type;
        if (t == GOAL) {
            c.h->reward += GOAL_REWARD;
            c.h->level_complete = 1;
            c.h->done = 1;
        } else if (t == OBSTACLE || t == ENEMY || t == TARGET) {
            c.h->done = 1;
        }
    }
    // caveflyer.cpp:72-79
    static PG_HD void update_agent_velocity(Ctx &c) {
        EnvHdr &h = *c.h;
        Entity &a = agent_of(c);
        float v_scale = get_agent_acceleration_scale(c);
        a.vx = (float)((double)a.vx + (double)(h.mixrate * h.maxspeed * h.action_vx * v_scale) * .2);
        a.vy = (float)((double)a.vy + (double)(h.mixrate * h.maxspeed * h.action_vy * v_scale) * .2);
        E::decay_agent_velocity(c);
    }
    // caveflyer.cpp:85-92
    static PG_HD bool is_blocked(Ctx &c, int src, int target, bool is_horizontal) {
        if (Defaults<CaveFlyerGame>::is_blocked(c, src, target, is_horizontal))
            return true;
        if (c.ents[src].type == PLAYER && target == CAVEWALL)
            return true;
        return false;
    }
    // caveflyer.cpp:94-123
    static PG_HD void handle_collision(Ctx &c, int si, int ti) {
        if (c.ents[ti].type == PLAYER_BULLET) {
            bool erase_bullet = false;
            Entity &src = c.ents[si];
            if (src.type == TARGET) {
                src.health -= 1;
                erase_bullet = true;
                if (src.health <= 0 && !src.will_erase) {
                    E::spawn_child(c, si, EXPLOSION, (float)(.5 * c.ents[si].rx));
                    c.ents[si].will_erase = 1;
                    c.h->reward += TARGET_REWARD;
                }
            } else if (src.type == OBSTACLE || src.type == ENEMY || src.type == GOAL) {
                erase_bullet = true;
            }
            if (erase_bullet && !c.ents[ti].will_erase) {
                c.ents[ti].will_erase = 1;
                int xi = E::spawn_child(c, ti, EXPLOSION, (float)(.5 * c.ents[ti].rx));
                c.ents[xi].vx = c.ents[si].vx;
                c.ents[xi].vy = c.ents[si].vy;
            }
        }
    }
    static PG_HD bool will_reflect(Ctx &c, int src, int target) { return (src == ENEMY && (target == CAVEWALL || target == c.oob)); }
    // caveflyer.cpp:129-145
    static PG_HD void choose_world_dim(Ctx &c) {
        int dist_diff = c.h->options.distribution_mode;
        int world_dim = 20;
        if (dist_diff == EasyMode)
            world_dim = 30;
        else if (dist_diff == HardMode)
            world_dim = 40;
        else if (dist_diff == MemoryMode)
            world_dim = 60;
        c.h->main_width = world_dim;
        c.h->main_height = world_dim;
    }
    static PG_HD void simple_choose(Ctx &c, int n, int k, int32_t *chosen, int32_t *flags) {
        pg_warp_for(n, [=](int i) { flags[i] = 0; });
        for (int i = 0; i < k; i++) {
            int next = rand_randn(*c.rng, n);
            while (flags[next]) next = rand_randn(*c.rng, n);
            chosen[i] = next;
            flags[next] = 1;
        }
    }
    // caveflyer.cpp:147-266
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        MT19937 &rg = *c.rng;
        h.out_of_bounds_object = WALL_OBJ;
        ctx_refresh(c);
        const int n = h.grid_size;
        {
            // one rand01() per cell, in cell order (caveflyer.cpp:154-160): drawn in bulk, then thresholded
            uint32_t *raw = reinterpret_cast<uint32_t *>(c.scratch);
            rand_fill_raw(rg, raw, n);
            int16_t *g0 = c.grid;
            pg_warp_for(n, [=](int i) { g0[i] = (int16_t)((float)((double)raw[i] / 4294967296.0) < .5 ? WALL_OBJ : SPACE); });
        }
        RoomGen<CaveFlyerGame> rm;
        rm.init(c, c.scratch, 12 * GRID_CAP);
        int32_t *best_room = c.scratch + 12 * GRID_CAP;
        int32_t *free_cells = c.scratch + 13 * GRID_CAP;
        int32_t *goal_path = c.scratch + 14 * GRID_CAP;
        int32_t *wide_path = c.scratch + 15 * GRID_CAP;
        int32_t *flags = c.scratch + 16 * GRID_CAP;
        int32_t *chosen = c.scratch + 17 * GRID_CAP;
        if (!rm.ok)
            return;
        for (int iteration = 0; iteration < 4; iteration++) rm.update();
        int best_size = rm.find_best_room(best_room);
        if (best_size <= 0) {
            h.err |= ERR_FASSERT;
            return;
        }
        int16_t *g = c.grid;
        pg_warp_for(n, [=](int i) { g[i] = (int16_t)WALL_OBJ; });
        pg_warp_for(n, [=](int i) {
            if (best_room[i])
                g[i] = (int16_t)SPACE;
        });
        int nfree = pg_warp_compact(n, free_cells, [=](int i) { return best_room[i] != 0; });
        simple_choose(c, nfree, 2, chosen, flags);
        int agent_cell = free_cells[chosen[0]];
        int goal_cell = free_cells[chosen[1]];
        agent_of(c).x = (float)((agent_cell % h.main_width) + .5);
        agent_of(c).y = (float)((agent_cell / h.main_width) + .5);
        int gi = E::spawn_entity_at_idx(c, goal_cell, .5, GOAL);
        c.ents[gi].collides_with_entities = 1;
        int path_len = rm.find_path(agent_cell, goal_cell, goal_path);
        bool should_prune = h.options.distribution_mode != MemoryMode;
        if (should_prune) {
            pg_warp_for(n, [=](int i) { wide_path[i] = 0; });
            for (int q = 0; q < path_len; q++) wide_path[goal_path[q]] = 1;
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
            rm.expand_room(wide_path, 4);
            pg_warp_for(n, [=](int i) { g[i] = (int16_t)(wide_path[i] ? SPACE : WALL_OBJ); });
        }
        for (int iteration = 0; iteration < 4; iteration++) {
            rm.update();
            for (int q = 0; q < path_len; q++) c.grid[goal_path[q]] = (int16_t)SPACE;
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
        }
        for (int q = 0; q < path_len; q++) c.grid[goal_path[q]] = (int16_t)MARKER;
        nfree = pg_warp_compact(n, free_cells, [=](int i) { return g[i] == SPACE; });
        pg_warp_for(n, [=](int i) {
            if (g[i] == WALL_OBJ)
                g[i] = (int16_t)CAVEWALL;
        });
        int chunk_size = nfree / 80;
        int num_objs = 3 * chunk_size;
        if (num_objs > GRID_CAP) {
            h.err |= ERR_SCRATCH_OVERFLOW;
            num_objs = 0;
        }
        simple_choose(c, nfree, num_objs, chosen, flags);
        for (int i = 0; i < num_objs; i++) {
            int val = free_cells[chosen[i]];
            if (i < chunk_size) {
                int ei = E::spawn_entity_at_idx(c, val, .5, OBSTACLE);
                c.ents[ei].collides_with_entities = 1;
            } else if (i < 2 * chunk_size) {
                int ei = E::spawn_entity_at_idx(c, val, .5, TARGET);
                c.ents[ei].health = 5;
                c.ents[ei].collides_with_entities = 1;
            } else {
                int ei = E::spawn_entity_at_idx(c, val, .5, ENEMY);
                // `(.1 * rand01() + .1) * (randn(2) * 2 - 1)`: operand order as compiled by g++ (left first)
                double mag = .1 * (double)rand_rand01(rg) + .1;
                float vel = (float)(mag * (rand_randn(rg, 2) * 2 - 1));
                if (rand_rand01(rg) < .5)
                    c.ents[ei].vx = vel;
                else
                    c.ents[ei].vy = vel;
                c.ents[ei].smart_step = 1;
                c.ents[ei].collides_with_entities = 1;
            }
        }
        pg_warp_for(n, [=](int i) {
            if (g[i] == MARKER)
                g[i] = (int16_t)SPACE;
        });
        h.out_of_bounds_object = CAVEWALL;
        ctx_refresh(c);
        h.visibility = h.options.distribution_mode == EasyMode ? 10 : 16;
    }
    // caveflyer.cpp:268-287 — sin/cos are the double overloads
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        EnvHdr &h = *c.h;
        Entity &a = agent_of(c);
        float acceleration = move_action % 3 - 1;
        if (acceleration < 0)
            acceleration *= 0.33f;
        float theta = -1 * a.rotation + PI_F / 2;
        if (acceleration > 0) {
            float ex = (float)((double)a.x - (double)a.rx * cos((double)theta));
            float ey = (float)((double)a.y - (double)a.ry * sin((double)theta));
            int xi = E::add_entity(c, ex, ey, 0, 0, (float)(.5 * agent_of(c).rx), EXHAUST);
            Entity &ex_ent = c.ents[xi];
            ex_ent.expire_time = 4;
            ex_ent.rotation = -1 * theta - PI_F / 2;
            ex_ent.grow_rate = 1.25;
            ex_ent.alpha_decay = 0.8f;
        }
        h.action_vy = (float)((double)acceleration * sin((double)theta));
        h.action_vx = (float)((double)acceleration * cos((double)theta));
        h.action_vrot = move_action / 3 - 1;
    }
    // caveflyer.cpp:289-325
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        if (h.special_action == 1) {
            Entity &a = agent_of(c);
            float theta = -1 * a.rotation + PI_F / 2;
            float vx = (float)cos((double)theta);
            float vy = (float)sin((double)theta);
            int bi = E::add_entity_rxy(c, a.x, a.y, vx, vy, 0.1f, 0.25f, PLAYER_BULLET);
            c.ents[bi].expire_time = 10;
            c.ents[bi].rotation = agent_of(c).rotation;
        }
        for (int ent_idx = h.n_ents - 1; ent_idx >= 0; ent_idx--) {
            Entity &ent = c.ents[ent_idx];
            if (ent.type == ENEMY) {
                entity_face_direction(ent, ent.vx, ent.vy, -1 * PI_F / 2);
            }
            if (ent.type != PLAYER_BULLET)
                continue;
            bool found_wall = false;
            for (int i = 0; i < 2; i++) {
                for (int j = 0; j < 2; j++) {
                    int type2 = E::get_obj_from_floats(c, ent.x + ent.rx * (2 * i - 1), ent.y + ent.ry * (2 * j - 1));
                    found_wall = found_wall || type2 == CAVEWALL;
                }
            }
            if (found_wall) {
                ent.will_erase = 1;
                E::spawn_child(c, ent_idx, EXPLOSION, (float)(.5 * c.ents[ent_idx].rx));
            }
        }
        E::erase_if_needed(c);
    }
};

}  // namespace pg
