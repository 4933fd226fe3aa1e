// CoinRun on the device engine. Behaviour restated from games/coinrun.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct CoinRunState {
    float last_agent_y;
    int32_t wall_theme;
    int32_t has_support;
    int32_t facing_right;
    int32_t is_on_crate;
    float gravity;
    float air_control;
};

struct CoinRun : Defaults<CoinRun>, DrawDefaults<CoinRun> {
    using E = Engine<CoinRun>;
    static constexpr int ENT_CAP = 512;  // observed > 256 live entities (trails of many enemies) in 65 536-env rollouts
    static constexpr int GRID_CAP = 64 * 64;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 512;  // = ENT_CAP: the blit list lives in global memory, no reason to be tight
    static constexpr int MAX_ROT_BLITS = 0;
    static constexpr int MAX_VIEW_CELLS = 17;  // visibility 13: int(c-7.5)..int(c+7.5) -> <= 16 cells
    static constexpr int FULL_VIEW_CELLS = 64;  // center_agent = false: the whole world (basic-abstract-game.cpp:819-838)
    static constexpr const char *NAME = "coinrun";

    // coinrun.cpp:11-34
    static constexpr float GOAL_REWARD = 10.0f;
    static constexpr int GOAL = 1, SAW = 2, SAW2 = 3, ENEMY = 5, ENEMY1 = 6, ENEMY2 = 7;
    static constexpr int PLAYER_JUMP = 9, PLAYER_RIGHT1 = 12, PLAYER_RIGHT2 = 13;
    static constexpr int WALL_MID = 15, WALL_TOP = 16, LAVA_MID = 17, LAVA_TOP = 18, ENEMY_BARRIER = 19;
    static constexpr int CRATE = 20;
    static constexpr int NUM_GROUND_THEMES = 6;

    static PG_HD CoinRunState &st(Ctx &c) { return game_state<CoinRunState>(c); }
    static PG_HD bool is_wall(int type) { return type == WALL_MID || type == WALL_TOP; }
    static PG_HD bool is_lava(int type) { return type == LAVA_MID || type == LAVA_TOP; }

    // coinrun.cpp:49-57
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->visibility = 13;
        c.h->mixrate = 0.2f;
        c.h->main_width = 64;
        c.h->main_height = 64;
        c.h->out_of_bounds_object = WALL_MID;
    }

    // coinrun.cpp:123-131
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        int t = c.ents[oi].type;
        if (t == ENEMY)
            c.h->done = 1;
        else if (t == SAW)
            c.h->done = 1;
    }
    // coinrun.cpp:133-138
    static PG_HD int theme_for_grid_obj(Ctx &c, int type) {
        if (is_wall(type))
            return st(c).wall_theme;
        return 0;
    }
    // coinrun.cpp:140-142
    static PG_HD bool will_reflect(Ctx &c, int src, int target) {
        return (src == ENEMY && (is_wall(target) || target == ENEMY_BARRIER));
    }
    // is_blocked_ents (:187-203) can only be true for CRATE or a type is_blocked accepts; will_reflect
    // (:140-142) only for wall-like types. Entity types here: PLAYER, SAW, ENEMY, CRATE, TRAIL.
    static PG_HD bool may_be_obstacle(Ctx &c, int target) { return may_block_or_reflect(c, 0, target); }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int target) {
        return target == CRATE || target == WALL_OBJ || target == c.oob || is_wall(target) || target == ENEMY_BARRIER;
    }
    // coinrun.cpp:144-154
    static PG_HD void handle_grid_collision(Ctx &c, int oi, int type, int i, int j) {
        if (c.ents[oi].type == PLAYER) {
            if (type == GOAL) {
                c.h->reward += GOAL_REWARD;
                c.h->done = 1;
                c.h->level_complete = 1;
            } else if (is_lava(type)) {
                c.h->done = 1;
            }
        }
    }
    // coinrun.cpp:156-173
    static PG_HD void update_agent_velocity(Ctx &c) {
        EnvHdr &h = *c.h;
        CoinRunState &s = st(c);
        Entity &a = agent_of(c);
        float mixrate_x = s.has_support ? h.mixrate : (h.mixrate * s.air_control);
        a.vx = (1 - mixrate_x) * a.vx + mixrate_x * h.maxspeed * h.action_vx;
        if (pg_dfabs((double)a.vx) < (double)(mixrate_x * h.maxspeed))
            a.vx = 0;
        if (h.action_vy > 0) {
            a.vy = h.max_jump;
        } else {
            if (s.has_support)
                a.vy = (float)((double)a.vy + .2 * (double)h.action_vy);
        }
        if (!(s.has_support && h.action_vy > 0)) {
            a.vy -= s.gravity;
            a.vy = pg_clip_abs(a.vy, h.max_jump);
        }
    }
    // coinrun.cpp:187-211
    static PG_HD bool is_blocked_ents(Ctx &c, int src, int target, bool is_horizontal) {
        Entity &t = c.ents[target];
        if (t.type == CRATE && !is_horizontal) {
            Entity &a = agent_of(c);
            if (a.vy >= 0)
                return false;
            if (c.h->action_vy < 0)
                return false;
            if (st(c).last_agent_y < (t.y + t.ry + a.ry))
                return false;
            st(c).is_on_crate = 1;
            return true;
        }
        return is_blocked(c, src, t.type, is_horizontal);
    }
    static PG_HD bool is_blocked(Ctx &c, int src, int target, bool is_horizontal) {
        if (Defaults<CoinRun>::is_blocked(c, src, target, is_horizontal))
            return true;
        if (c.ents[src].type == PLAYER && is_wall(target))
            return true;
        return false;
    }
    // coinrun.cpp:213-225
    static PG_HD int image_for_type(Ctx &c, int type) {
        if (type == PLAYER) {
            if (pg_dfabs((double)agent_of(c).vx) < .01 && c.h->action_vx == 0 && st(c).has_support)
                return PLAYER;
            return (c.h->cur_time / 5 % 2 == 0 || !st(c).has_support) ? PLAYER_RIGHT1 : PLAYER_RIGHT2;
        } else if (type == ENEMY_BARRIER) {
            return -1;
        }
        return Defaults<CoinRun>::image_for_type(c, type);
    }
    // coinrun.cpp:64-70
    static PG_HD bool get_adjusted_image_rect(Ctx &c, int type, double *adj) {
        if (type == PLAYER || type == PLAYER_JUMP || type == PLAYER_RIGHT1 || type == PLAYER_RIGHT2) {
            adj[0] = 0;
            adj[1] = -.7415;
            adj[2] = 1;
            adj[3] = 1.7415;
            return true;
        }
        return false;
    }

    // ---- level generation (coinrun.cpp:227-414)
    static PG_HD void fill_block_top(Ctx &c, int x, int y, int dx, int dy, int fill, int top) {
        E::fill_elem(c, x, y, dx, dy - 1, fill);
        E::fill_elem(c, x, y + dy - 1, dx, 1, top);
    }
    static PG_HD void fill_ground_block(Ctx &c, int x, int y, int dx, int dy) { fill_block_top(c, x, y, dx, dy, WALL_MID, WALL_TOP); }
    static PG_HD void fill_lava_block(Ctx &c, int x, int y, int dx, int dy) { fill_block_top(c, x, y, dx, dy, LAVA_MID, LAVA_TOP); }
    static PG_HD void init_floor_and_walls(Ctx &c) {
        int w = c.h->main_width, h = c.h->main_height;
        E::fill_elem(c, 0, 0, w, 1, WALL_TOP);
        E::fill_elem(c, 0, 0, 1, h, WALL_MID);
        E::fill_elem(c, w - 1, 0, 1, h, WALL_MID);
        E::fill_elem(c, 0, h - 1, w, 1, WALL_MID);
    }
    static PG_HD void create_saw_enemy(Ctx &c, int x, int y) { E::add_entity(c, (float)(x + .5), (float)(y + .5), 0, 0, .5, SAW); }
    static PG_HD void create_enemy(Ctx &c, int x, int y) {
        float vx = (float)(.15 * (rand_randn(*c.rng, 2) * 2 - 1));
        int ei = E::add_entity(c, (float)(x + .5), (float)(y + .5), vx, 0, .5, ENEMY);
        Entity &ent = c.ents[ei];
        ent.smart_step = 1;
        ent.image_type = ENEMY1;
        ent.render_z = 1;
        E::choose_random_theme(c, ent);
    }
    static PG_HD void create_crate(Ctx &c, int x, int y) {
        int ei = E::add_entity(c, (float)(x + .5), (float)(y + .5), 0, 0, .5, CRATE);
        E::choose_random_theme(c, c.ents[ei]);
    }

    static PG_HD void generate_coin_to_the_right(Ctx &c) {
        EnvHdr &h = *c.h;
        MT19937 &rg = *c.rng;
        int max_difficulty = 3;
        int dif = rand_randn(rg, max_difficulty) + 1;
        int num_sections = rand_randn(rg, dif) + dif;
        int curr_x = 5;
        int curr_y = 1;
        int pit_threshold = dif;
        int danger_type = rand_randn(rg, 3);
        bool allow_pit = (h.options.debug_mode & (1 << 1)) == 0;
        bool allow_crate = (h.options.debug_mode & (1 << 2)) == 0;
        bool allow_dy = (h.options.debug_mode & (1 << 3)) == 0;
        int w = h.main_width;
        float gravity = st(c).gravity;
        float _max_dy = h.max_jump * h.max_jump / (2 * gravity);
        float _max_dx = h.maxspeed * 2 * h.max_jump / gravity;
        int max_dy = (int)(_max_dy - .5);
        int max_dx = (int)(_max_dx - .5);
        bool allow_monsters = true;
        if (h.options.distribution_mode == EasyMode)
            allow_monsters = false;

        for (int section_idx = 0; section_idx < num_sections; section_idx++) {
            if (curr_x + 15 >= w)
                break;
            int dy = rand_randn(rg, 4) + 1 + int(dif / 3);
            if (!allow_dy)
                dy = 0;
            if (dy > max_dy)
                dy = max_dy;
            if (curr_y >= 20) {
                dy *= -1;
            } else if (curr_y >= 5 && rand_randn(rg, 2) == 1) {
                dy *= -1;
            }
            int dx = rand_randn(rg, 2 * dif) + 3 + int(dif / 3);
            curr_y += dy;
            if (curr_y < 1)
                curr_y = 1;
            bool use_pit = allow_pit && (dx > 7) && (curr_y > 3) && (rand_randn(rg, 20) >= pit_threshold);
            if (use_pit) {
                int x1 = rand_randn(rg, 3) + 1;
                int x2 = rand_randn(rg, 3) + 1;
                int pit_width = dx - x1 - x2;
                if (pit_width > max_dx) {
                    pit_width = max_dx;
                    x2 = dx - x1 - pit_width;
                }
                fill_ground_block(c, curr_x, 0, x1, curr_y);
                fill_ground_block(c, curr_x + dx - x2, 0, x2, curr_y);
                int lava_height = rand_randn(rg, curr_y - 3) + 1;
                if (danger_type == 0) {
                    fill_lava_block(c, curr_x + x1, 1, pit_width, lava_height);
                } else if (danger_type == 1) {
                    for (int ei = 0; ei < pit_width; ei++) create_saw_enemy(c, curr_x + x1 + ei, 1);
                } else if (danger_type == 2) {
                    for (int ei = 0; ei < pit_width; ei++) create_enemy(c, curr_x + x1 + ei, 1);
                }
                if (pit_width > 4) {
                    int x3, w1;
                    if (pit_width == 5) {
                        x3 = 1 + rand_randn(rg, 2);
                        w1 = 1 + rand_randn(rg, 2);
                    } else if (pit_width == 6) {
                        x3 = 2 + rand_randn(rg, 2);
                        w1 = 1 + rand_randn(rg, 2);
                    } else {
                        x3 = 2 + rand_randn(rg, 2);
                        int x4 = 2 + rand_randn(rg, 2);
                        w1 = pit_width - x3 - x4;
                    }
                    fill_ground_block(c, curr_x + x1 + x3, curr_y - 1, w1, 1);
                }
            } else {
                fill_ground_block(c, curr_x, 0, dx, curr_y);
                int ob1_x = -1;
                int ob2_x = -1;
                if (rand_randn(rg, 10) < (2 * dif) && dx > 3) {
                    ob1_x = curr_x + rand_randn(rg, dx - 2) + 1;
                    create_saw_enemy(c, ob1_x, curr_y);
                }
                if (rand_randn(rg, 10) < dif && dx > 3 && (max_dx >= 4) && allow_monsters) {
                    ob2_x = curr_x + rand_randn(rg, dx - 2) + 1;
                    create_enemy(c, ob2_x, curr_y);
                }
                if (allow_crate) {
                    for (int i = 0; i < 2; i++) {
                        int crate_x = curr_x + rand_randn(rg, dx - 2) + 1;
                        if (rand_randn(rg, 2) == 1 && ob1_x != crate_x && ob2_x != crate_x) {
                            int pile_height = rand_randn(rg, 3) + 1;
                            for (int j = 0; j < pile_height; j++) create_crate(c, crate_x, curr_y + j);
                        }
                    }
                }
            }
            if (!is_wall(E::get_obj(c, curr_x - 1, curr_y)))
                E::set_obj(c, curr_x - 1, curr_y, ENEMY_BARRIER);
            curr_x += dx;
            E::set_obj(c, curr_x, curr_y, ENEMY_BARRIER);
        }
        E::set_obj(c, curr_x, curr_y, GOAL);
        fill_ground_block(c, curr_x, 0, 1, curr_y);
        E::fill_elem(c, curr_x + 1, 0, h.main_width - curr_x - 1, h.main_height, WALL_MID);
    }

    // coinrun.cpp:416-445
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        CoinRunState &s = st(c);
        s.gravity = 0.2f;
        h.max_jump = 1.5;
        s.air_control = 0.15f;
        h.maxspeed = .5;
        s.has_support = 0;
        s.facing_right = 1;
        Entity &a = agent_of(c);
        if (h.options.distribution_mode == EasyMode) {
            a.image_theme = 0;
            s.wall_theme = 0;
            h.background_index = 0;
        } else {
            E::choose_random_theme(c, a);
            s.wall_theme = rand_randn(*c.rng, NUM_GROUND_THEMES);
        }
        a.rx = .5;
        a.ry = 0.5787f;
        a.x = 1 + a.rx;
        a.y = 1 + a.ry;
        s.last_agent_y = a.y;
        s.is_on_crate = 0;
        init_floor_and_walls(c);
        generate_coin_to_the_right(c);
    }

    static PG_HD bool can_support(Ctx &c, int obj) { return is_wall(obj) || obj == c.h->out_of_bounds_object; }

    // coinrun.cpp:451-472
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        EnvHdr &h = *c.h;
        CoinRunState &s = st(c);
        Entity &a = agent_of(c);
        h.action_vx = move_action / 3 - 1;
        h.action_vy = (move_action % 3) - 1;
        if (h.action_vx > 0)
            s.facing_right = 1;
        if (h.action_vx < 0)
            s.facing_right = 0;
        float yb = (float)((double)a.y - ((double)a.ry + .01));
        int obj_below_1 = E::get_obj_from_floats(c, (float)((double)a.x - ((double)a.rx - .01)), yb);
        int obj_below_2 = E::get_obj_from_floats(c, (float)((double)a.x + ((double)a.rx - .01)), yb);
        s.has_support = (s.is_on_crate || can_support(c, obj_below_1) || can_support(c, obj_below_2)) && a.vy == 0;
        s.is_on_crate = 0;
        if (h.action_vy == 1) {
            if (!s.has_support)
                h.action_vy = 0;
        }
    }

    // coinrun.cpp:474-498
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        if (h.action_vx > 0)
            agent_of(c).is_reflected = 0;
        if (h.action_vx < 0)
            agent_of(c).is_reflected = 1;
        for (int i = h.n_ents - 1; i >= 0; i--) {
            Entity &ent = c.ents[i];
            if (ent.type == ENEMY) {
                int ti = E::add_entity_rxy(c, ent.x, (float)((double)ent.y - (double)ent.ry * .5), 0, 0.01f, 0.3f, 0.2f, TRAIL);
                c.ents[ti].expire_time = 8;
                c.ents[ti].alpha = .5;
                ent.image_type = h.cur_time / 5 % 2 == 0 ? ENEMY1 : ENEMY2;
                ent.is_reflected = ent.vx > 0;
            } else if (ent.type == SAW) {
                ent.image_type = h.cur_time % 2 == 0 ? SAW : SAW2;
            }
        }
        st(c).last_agent_y = agent_of(c).y;
    }
};

}  // namespace pg
