// Jumper on the device engine. Behaviour restated from games/jumper.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"
#include "../pg_mazegen.cuh"
#include "../pg_roomgen.cuh"

namespace pg {

struct JumperState {
    int32_t goal_idx;  // the reference holds a shared_ptr to the goal entity
    int32_t jump_count, jump_delta, jump_time, has_support, facing_right, wall_theme;
    float compass_dim;
};

struct JumperGame : Defaults<JumperGame>, DrawDefaults<JumperGame> {
    using E = Engine<JumperGame>;
    static constexpr int ENT_CAP = 320;  // goal + up to ~20% of the floor cells as spikes + trails
    static constexpr int GRID_CAP = 45 * 45;
    static constexpr int MAZE_WORDS = 3200;  // MazeGen::words_needed(15) = 3081
    static constexpr int ROOM_WORDS = 12 * GRID_CAP;
    static constexpr int SCRATCH_WORDS = MAZE_WORDS + ROOM_WORDS + 5 * GRID_CAP;
    static constexpr int MAX_VISIBLE_ENTS = 320;
    static constexpr int MAX_ROT_BLITS = 3;    // no sprite rotates; the slots hold the compass disc, its needle and the double-jump shadow
    static constexpr int MAX_VIEW_CELLS = 20;  // visibility 16: int(c-9)..int(c+9)
    static constexpr int FULL_VIEW_CELLS = 45;  // center_agent = false: the whole world (basic-abstract-game.cpp:819-838)
    static constexpr const char *NAME = "jumper";

    // jumper.cpp:11-27
    static constexpr float GOAL_REWARD = 10.0f;
    static constexpr int GOAL = 1, SPIKE = 2, CAVEWALL = 6, CAVEWALL_TOP = 7;
    static constexpr int PLAYER_JUMP = 9, PLAYER_LEFT1 = 10, PLAYER_LEFT2 = 11, PLAYER_RIGHT1 = 12, PLAYER_RIGHT2 = 13;
    static constexpr int MAZE_SCALE = 3, JUMP_COOLDOWN = 3, NUM_WALL_THEMES = 4;

    static PG_HD JumperState &st(Ctx &c) { return game_state<JumperState>(c); }
    static PG_HD bool is_wall(int obj) { return obj == CAVEWALL || obj == CAVEWALL_TOP; }
    static PG_HD bool can_support(Ctx &c, int obj) { return is_wall(obj) || obj == c.h->out_of_bounds_object; }

    static constexpr bool HAS_ENTITY_HOOKS = true;
    static PG_HD void on_entity_moved(Ctx &c, int from, int to) {
        if (st(c).goal_idx == from)
            st(c).goal_idx = to;
    }
    // jumper.cpp:80-90
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        int t = c.ents[oi].type;
        if (t == GOAL) {
            c.h->reward += GOAL_REWARD;
            c.h->level_complete = 1;
            c.h->done = 1;
        } else if (t == SPIKE) {
            c.h->done = 1;
        }
    }
    // jumper.cpp:92-98
    static PG_HD void update_agent_velocity(Ctx &c) {
        EnvHdr &h = *c.h;
        Entity &a = agent_of(c);
        float v_scale = get_agent_acceleration_scale(c);
        a.vx = (1 - h.mixrate) * a.vx + h.mixrate * h.maxspeed * h.action_vx * v_scale;
        if (h.action_vy != 0)
            a.vy = h.maxspeed * h.action_vy * 2;
    }
    // jumper.cpp:100-105
    static PG_HD int theme_for_grid_obj(Ctx &c, int type) { return is_wall(type) ? st(c).wall_theme : 0; }
    // jumper.cpp:111-118
    static PG_HD bool is_blocked(Ctx &c, int src, int target, bool is_horizontal) {
        if (Defaults<JumperGame>::is_blocked(c, src, target, is_horizontal))
            return true;
        if (c.ents[src].type == PLAYER && is_wall(target))
            return true;
        return false;
    }
    // no entity carries a wall id: entity overlaps never block or reflect (jumper.cpp:111-118, 177-179)
    static PG_HD bool may_be_obstacle(Ctx &c, int target) { return false; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int target) { return false; }
    // jumper.cpp:120-135
    static PG_HD int image_for_type(Ctx &c, int type) {
        if (type == PLAYER) {
            EnvHdr &h = *c.h;
            JumperState &s = st(c);
            if (pg_dfabs((double)agent_of(c).vx) < .01 && h.action_vx == 0 && s.has_support)
                return PLAYER;
            const bool first = h.cur_time / 5 % 2 == 0 || !s.has_support;
            if (s.facing_right)
                return first ? PLAYER_RIGHT1 : PLAYER_RIGHT2;
            return first ? PLAYER_LEFT1 : PLAYER_LEFT2;
        }
        return type < 0 ? -type : type;
    }
    // jumper.cpp:137-177: the compass (disc, needle, distance bar) and the double-jump shadow
    template <class Frame>
    static PG_HD void make_overlay_blits(Ctx &c, Frame &f) {
        EnvHdr &h = *c.h;
        JumperState &s = st(c);
        if (h.options.distribution_mode == MemoryMode)
            return;
        const Entity &a = agent_of(c);
        const Entity &goal = c.ents[s.goal_idx];
        double cr_[4];
        Raster<JumperGame, Frame>::abs_rect(f.cam, (float)((double)(h.view_dim - s.compass_dim) - .25), .25, s.compass_dim, s.compass_dim, cr_);
        const uint32_t clock_color = (168u << 16) | (166u << 8) | 158u;
        const uint32_t highlight = (252u << 16) | (186u << 8) | 3u;
        int n = f.n_overlay;
        bool ok = make_ellipse_blit(f, f.overlay[n++], 0, cr_[0], cr_[1], cr_[2], cr_[3], 0xff000000u | clock_color, true);
        // QRectF::center() = x + w/2 in double, narrowed to float (jumper.cpp:146-148)
        float cx = (float)(cr_[0] + cr_[2] / 2);
        float cy = (float)(cr_[1] + cr_[3] / 2);
        float cr = (float)(cr_[2] / 2 * .95);
        // get_theta (basic-abstract-game.cpp:233-238): float differences, the C double atan2 (checked
        // in the oracle's object code — unlike entity.cpp this file does not pull in the float overloads)
        float theta = (float)atan2((double)(goal.y - a.y), (double)(goal.x - a.x));
        // QPainter::drawLine(int, int, int, int): the four values are truncated to int
        const int lx1 = (int)cx, ly1 = (int)cy;
        const int lx2 = (int)((double)cx + (double)cr * cos((double)theta));
        const int ly2 = (int)((double)cy - (double)cr * sin((double)theta));
        ok = make_line_blit(f, f.overlay[n++], 1, lx1, ly1, lx2, ly2, 0xff000000u | highlight) && ok;
        float dist = E::get_distance(a, goal);
        float dist_pct = (float)((double)dist / ((double)h.main_width * pg_dsqrt(2.0)));
        float bar_thickness = s.compass_dim / 8;
        double br[4];
        Raster<JumperGame, Frame>::abs_rect(f.cam, (float)((double)(h.view_dim - s.compass_dim) - .25), (float)(.25 + (double)s.compass_dim),
                                            s.compass_dim * dist_pct, bar_thickness, br);
        make_solid_blit(f.overlay[n++], br[0], br[1], br[2], br[3], highlight);
        if (s.jump_delta < 0 && !s.has_support) {
            double r1[4];
            Raster<JumperGame, Frame>::object_rect(f.cam, a, r1);
            // QRect(int, int, int, int) from doubles: truncation; white at alpha 120, no pen
            const int ex = (int)r1[0], ey = (int)(r1[1] + r1[3] * (5.0 / 6)), ew = (int)r1[2], eh = (int)(r1[3] / 3);
            // QColor(255, 255, 255, 120) premultiplied
            ok = make_ellipse_blit(f, f.overlay[n++], 2, (double)ex, (double)ey, (double)ew, (double)eh, 0x78787878u, false) && ok;
        }
        if (!ok)
            h.err |= ERR_UNSUPPORTED;
        f.n_overlay = n;
    }
    // jumper.cpp:181-199
    static PG_HD bool is_space_on_ground(Ctx &c, int x, int y) {
        if (E::get_obj(c, x, y) != SPACE)
            return false;
        if (E::get_obj(c, x, y + 1) != SPACE)
            return false;
        int below_obj = E::get_obj(c, x, y - 1);
        return below_obj == CAVEWALL || below_obj == c.oob;
    }
    static PG_HD bool is_top_wall(Ctx &c, int x, int y) { return E::get_obj(c, x, y) == CAVEWALL && E::get_obj(c, x, y + 1) == SPACE; }
    static PG_HD bool is_left_wall(Ctx &c, int x, int y) { return E::get_obj(c, x, y) == CAVEWALL && E::get_obj(c, x + 1, y) == SPACE; }
    static PG_HD bool is_right_wall(Ctx &c, int x, int y) { return E::get_obj(c, x, y) == CAVEWALL && E::get_obj(c, x - 1, y) == SPACE; }
    // jumper.cpp:201-217
    static PG_HD void choose_world_dim(Ctx &c) {
        int dist_diff = c.h->options.distribution_mode;
        int world_dim = 20;
        if (dist_diff == EasyMode)
            world_dim = 20;
        else if (dist_diff == HardMode)
            world_dim = 40;
        else if (dist_diff == MemoryMode)
            world_dim = 45;
        c.h->main_width = world_dim;
        c.h->main_height = world_dim;
    }
    // jumper.cpp:219-368
    static PG_HD void game_reset(Ctx &c) {
        EnvHdr &h = *c.h;
        JumperState &s = st(c);
        MT19937 &rg = *c.rng;
        if (h.options.distribution_mode == EasyMode) {
            h.visibility = 12;
            s.compass_dim = 3;
        } else {
            h.visibility = 16;
            s.compass_dim = 2;
        }
        if (h.options.distribution_mode == MemoryMode)
            h.timeout = 2000;
        PG_PHASE_RESET(c);
        PG_PHASE_BEGIN(c);
        E::basic_game_reset(c);
        PG_PHASE_END(c, 0);
        h.out_of_bounds_object = WALL_OBJ;
        ctx_refresh(c);
        s.wall_theme = rand_randn(rg, NUM_WALL_THEMES);
        s.jump_count = 0;
        s.jump_delta = 0;
        s.jump_time = 0;
        s.has_support = 0;
        s.facing_right = 1;
        s.goal_idx = 0;
        const int n = h.grid_size;
        const int w = h.main_width;
        {
            const int maze_dim = w / MAZE_SCALE;
            MazeGen mg;
            mg.init(c, maze_dim);
            mg.generate_maze_no_dead_ends();
            PG_PHASE_END(c, 1);
            // one rand01() per cell, in cell order (jumper.cpp:246-255): drawn in bulk behind the
            // maze workspace, then thresholded against the maze cell's probability
            uint32_t *raw = reinterpret_cast<uint32_t *>(c.scratch + MAZE_WORDS);
            rand_fill_raw(rg, raw, n);
            int16_t *g0 = c.grid;
            const MazeGen *mgp = &mg;
            pg_warp_for(n, [=](int i) {
                int obj = mgp->grid_get((i % w) / MAZE_SCALE + 1, (i / w) / MAZE_SCALE + 1);
                float prob = obj == WALL_OBJ ? .8 : .2;
                g0[i] = (int16_t)((float)((double)raw[i] / 4294967296.0) < prob ? WALL_OBJ : SPACE);
            });
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        PG_PHASE_END(c, 2);
        RoomGen<JumperGame> rm;
        rm.init(c, c.scratch + MAZE_WORDS, ROOM_WORDS);
        int32_t *best_room = c.scratch + MAZE_WORDS + ROOM_WORDS;
        int32_t *free_cells = best_room + GRID_CAP;
        int32_t *candidates = free_cells + GRID_CAP;
        int32_t *goal_path = candidates + GRID_CAP;
        int32_t *wide_path = goal_path + GRID_CAP;
        if (!rm.ok)
            return;
        for (int iteration = 0; iteration < 2; iteration++) rm.update();
        PG_PHASE_END(c, 3);
        // border cells (jumper.cpp:258-267)
        for (int i = 0; i < w; i++) {
            E::set_obj(c, i, 0, CAVEWALL);
            E::set_obj(c, i, h.main_height - 1, CAVEWALL);
        }
        for (int i = 0; i < h.main_height; i++) {
            E::set_obj(c, 0, i, CAVEWALL);
            E::set_obj(c, w - 1, i, CAVEWALL);
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        int best_size = rm.find_best_room(best_room);
        PG_PHASE_END(c, 4);
        if (best_size <= 0) {
            h.err |= ERR_FASSERT;
            return;
        }
        int16_t *g = c.grid;
        pg_warp_for(n, [=](int i) { g[i] = (int16_t)CAVEWALL; });
        pg_warp_for(n, [=](int i) {
            if (best_room[i])
                g[i] = (int16_t)SPACE;
        });
        const int nfree = pg_warp_compact(n, free_cells, [=](int i) { return best_room[i] != 0; });
        int goal_cell = free_cells[rand_randn(rg, nfree)];
        Ctx *cp = &c;
        const int ncand = pg_warp_compact(n, candidates, [=](int i) { return is_space_on_ground(*cp, i % w, i / w); });
        if (ncand <= 0) {
            h.err |= ERR_FASSERT;
            return;
        }
        int agent_cell = candidates[rand_randn(rg, ncand)];
        PG_PHASE_END(c, 5);
        int path_len = rm.find_path(agent_cell, goal_cell, goal_path);
        PG_PHASE_END(c, 6);
        bool should_prune = h.options.distribution_mode != MemoryMode;
        if (should_prune) {
            pg_warp_for(n, [=](int i) { wide_path[i] = 0; });
            for (int q = 0; q < path_len; q++) wide_path[goal_path[q]] = 1;
#if defined(__CUDA_ARCH__)
            __syncwarp();
#endif
            rm.expand_room(wide_path, 4);
            pg_warp_for(n, [=](int i) { g[i] = (int16_t)(wide_path[i] ? SPACE : CAVEWALL); });
        }
        PG_PHASE_END(c, 7);
        s.goal_idx = E::spawn_entity_at_idx(c, goal_cell, .5, GOAL);
        float spike_prob = h.options.distribution_mode == MemoryMode ? 0 : .2;
        // Both loops below walk the cells in order, and what they do at a cell (an RNG draw, a
        // changed cell) can change the test for later cells only. The warp finds the next cell
        // that passes the test, handles it exactly as the reference's loop body, and — if the grid
        // changed — re-tests from the next cell on.
        {
            ScanUpIter it(0, n);
            while (true) {
                const int i = it.next([=](int k) {
                    const int x = k % w, y = k / w;
                    return is_space_on_ground(*cp, x, y) && (is_space_on_ground(*cp, x - 1, y) && is_space_on_ground(*cp, x + 1, y));
                });
                if (i < 0)
                    break;
                if (rand_rand01(rg) < spike_prob) {
                    E::set_obj(c, i % w, i / w, SPIKE);
                    it.restart_from(i + 1);
                }
            }
        }
        // long vertical walls are broken up (jumper.cpp:323-335)
        {
            ScanUpIter it(0, n);
            while (true) {
                const int i = it.next([=](int k) {
                    const int x = k % w, y = k / w;
                    return (is_left_wall(*cp, x, y) && is_left_wall(*cp, x, y + 1) && is_left_wall(*cp, x, y + 2)) ||
                           (is_right_wall(*cp, x, y) && is_right_wall(*cp, x, y + 1) && is_right_wall(*cp, x, y + 2));
                });
                if (i < 0)
                    break;
                const int x = i % w, y = i / w;
                if (is_left_wall(c, x, y) && is_left_wall(c, x, y + 1) && is_left_wall(c, x, y + 2))
                    E::set_obj(c, x, y + rand_randn(rg, 3), SPACE);
                if (is_right_wall(c, x, y) && is_right_wall(c, x, y + 1) && is_right_wall(c, x, y + 2))
                    E::set_obj(c, x, y + rand_randn(rg, 3), SPACE);
                it.restart_from(i + 1);
            }
        }
        {
            Entity &a = agent_of(c);
            a.x = (float)((agent_cell % w) + .5);
            a.y = (agent_cell / w) + a.ry;
        }
        {
            // get_cells_with_type(SPIKE): ascending cell list (reusing the candidate buffer)
            const int nspikes = pg_warp_compact(n, candidates, [=](int i) { return g[i] == SPIKE; });
            for (int q = 0; q < nspikes; q++) {
                const int i = candidates[q];
                c.grid[i] = (int16_t)SPACE;
                float spike_ry = 0.4f;
                float spike_rx = 0.23f;
                E::add_entity_rxy(c, (float)((i % w) + .5), (i / w) + spike_ry, 0, 0, spike_rx, spike_ry, SPIKE);
            }
        }
#if defined(__CUDA_ARCH__)
        __syncwarp();
#endif
        // a wall cell with free space above gets the "top" sprite; turning a cell into a top wall
        // never changes another cell's test (it stays a non-SPACE cell), so cells are independent
        pg_warp_for(n, [=](int i) {
            if (is_top_wall(*cp, i % w, i / w))
                g[i] = (int16_t)CAVEWALL_TOP;
        });
        agent_of(c).rx = 0.254f;
        agent_of(c).ry = 0.4f;
        h.out_of_bounds_object = CAVEWALL;
        ctx_refresh(c);
        PG_PHASE_END(c, 8);
    }
    // jumper.cpp:378-423
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        EnvHdr &h = *c.h;
        JumperState &s = st(c);
        const Entity &a = agent_of(c);
        h.action_vx = move_action / 3 - 1;
        h.action_vy = (move_action % 3) - 1;
        if (h.action_vy < 0)
            h.action_vy = 0;
        if (h.action_vx > 0)
            s.facing_right = 1;
        if (h.action_vx < 0)
            s.facing_right = 0;
        float yb = (float)((double)a.y - ((double)a.ry + .01));
        int obj_below_1 = E::get_obj_from_floats(c, (float)((double)a.x - ((double)a.rx - .01)), yb);
        int obj_below_2 = E::get_obj_from_floats(c, (float)((double)a.x + ((double)a.rx - .01)), yb);
        s.jump_delta = 0;
        s.has_support = can_support(c, obj_below_1) || can_support(c, obj_below_2);
        if (s.has_support)
            s.jump_count = 2;
        if (h.action_vy == 1 && s.jump_count > 0 && (h.cur_time - s.jump_time > JUMP_COOLDOWN)) {
            s.jump_count -= 1;
            s.jump_delta = -1;
        } else {
            h.action_vy = 0;
        }
        if (h.action_vy > 0)
            s.jump_time = h.cur_time;
        h.action_vrot = 0;
    }
    // jumper.cpp:425-443
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        if (h.action_vx > 0)
            agent_of(c).is_reflected = 0;
        if (h.action_vx < 0)
            agent_of(c).is_reflected = 1;
        Entity &a = agent_of(c);
        if (pg_dfabs((double)a.vx) + pg_dfabs((double)a.vy) > .05) {
            float ax = a.x, ty = (float)((double)a.y - (double)a.ry * .5);
            int ti = E::add_entity_rxy(c, ax, ty, 0, 0.01f, 0.3f, 0.2f, TRAIL);
            c.ents[ti].expire_time = 8;
            c.ents[ti].alpha = .5;
        }
        if (agent_of(c).vy > -2)
            agent_of(c).vy -= 0.15f;
    }
};

}  // namespace pg
