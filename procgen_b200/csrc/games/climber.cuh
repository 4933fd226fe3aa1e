// Climber on the device engine. Behaviour restated from games/climber.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct ClimberState {
    int32_t has_support, facing_right, coin_quota, coins_collected, wall_theme;
    float gravity, air_control;
};

struct Climber : Defaults<Climber>, DrawDefaults<Climber> {
    using E = Engine<Climber>;
    static constexpr int ENT_CAP = 64;
    static constexpr int GRID_CAP = 20 * 64;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 64;
    static constexpr int MAX_ROT_BLITS = 0;
    static constexpr int MAX_VIEW_CELLS = 24;  // visibility = main_width (<= 20): int(c-11)..int(c+11)
    static constexpr int FULL_VIEW_CELLS = 64;  // center_agent = false: the whole world (basic-abstract-game.cpp:819-838)
    static constexpr const char *NAME = "climber";

    // climber.cpp:9-26
    static constexpr float COIN_REWARD = 1.0f;
    static constexpr float COMPLETION_BONUS = 10.0f;
    static constexpr int COIN = 1, ENEMY = 5, ENEMY1 = 6, ENEMY2 = 7, PLAYER_JUMP = 9, PLAYER_RIGHT1 = 12, PLAYER_RIGHT2 = 13;
    static constexpr int WALL_MID = 15, WALL_TOP = 16, ENEMY_BARRIER = 19;
    static constexpr float PATROL_RANGE = 4;
    static constexpr int NUM_WALL_THEMES = 4;

    static PG_HD ClimberState &st(Ctx &c) { return game_state<ClimberState>(c); }
    static PG_HD bool is_wall(int type) { return type == WALL_MID || type == WALL_TOP; }

    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->out_of_bounds_object = WALL_MID;
    }
    // climber.cpp:91-101
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        Entity &obj = c.ents[oi];
        if (obj.type == ENEMY) {
            c.h->done = 1;
        } else if (obj.type == COIN) {
            c.h->reward += COIN_REWARD;
            st(c).coins_collected += 1;
            obj.will_erase = 1;
        }
    }
    static PG_HD int theme_for_grid_obj(Ctx &c, int type) { return is_wall(type) ? st(c).wall_theme : 0; }
    static PG_HD bool will_reflect(Ctx &c, int src, int target) { return (src == ENEMY && (is_wall(target) || target == ENEMY_BARRIER)); }
    static PG_HD bool may_be_obstacle(Ctx &c, int target) { return may_block_or_reflect(c, 0, target); }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int target) {
        return target == WALL_OBJ || target == c.oob || is_wall(target) || target == ENEMY_BARRIER;
    }
    // climber.cpp:114-125
    static PG_HD void update_agent_velocity(Ctx &c) {
        EnvHdr &h = *c.h;
        ClimberState &s = st(c);
        Entity &a = agent_of(c);
        float mixrate_x = s.has_support ? h.mixrate : (h.mixrate * s.air_control);
        a.vx = (1 - mixrate_x) * a.vx + mixrate_x * h.maxspeed * h.action_vx;
        if (h.action_vy > 0)
            a.vy = h.max_jump;
        if (!s.has_support) {
            if (a.vy > -2)
                a.vy -= s.gravity;
        }
    }
    // climber.cpp:135-142
    static PG_HD bool is_blocked(Ctx &c, int src, int target, bool is_horizontal) {
        if (Defaults<Climber>::is_blocked(c, src, target, is_horizontal))
            return true;
        if (c.ents[src].type == PLAYER && is_wall(target))
            return true;
        return false;
    }
    // climber.cpp:144-158
    static PG_HD int image_for_type(Ctx &c, int type) {
        if (type == PLAYER) {
            if (!st(c).has_support)
                return PLAYER_JUMP;
            if (pg_dfabs((double)agent_of(c).vx) < .01 && c.h->action_vx == 0 && st(c).has_support)
                return PLAYER;
            return (c.h->cur_time / 5 % 2 == 0 || !st(c).has_support) ? PLAYER_RIGHT1 : PLAYER_RIGHT2;
        } else if (type == ENEMY_BARRIER) {
            return -1;
        }
        return Defaults<Climber>::image_for_type(c, type);
    }
    static PG_HD void init_floor_and_walls(Ctx &c) {
        int w = c.h->main_width, h = c.h->main_height;
        E::fill_elem(c, 0, 0, w, 1, WALL_TOP);
        E::fill_elem(c, 0, 0, 1, h, WALL_MID);
        E::fill_elem(c, w - 1, 0, 1, h, WALL_MID);
        E::fill_elem(c, 0, h - 1, w, 1, WALL_MID);
    }
    // climber.cpp:167-172
    static PG_HD int choose_delta_y(Ctx &c) {
        int max_dy = (int)(c.h->max_jump * c.h->max_jump / (2 * st(c).gravity));
        int min_dy = 3;
        return rand_randn(*c.rng, max_dy - min_dy + 1) + min_dy;
    }
    // climber.cpp:174-231
    static PG_HD void generate_platforms(Ctx &c) {
        EnvHdr &h = *c.h;
        ClimberState &s = st(c);
        MT19937 &rg = *c.rng;
        int difficulty = rand_randn(rg, 3);
        int min_platforms = difficulty * difficulty + 1;
        int max_platforms = (difficulty + 1) * (difficulty + 1) + 1;
        int num_platforms = rand_randn(rg, max_platforms - min_platforms + 1) + min_platforms;
        s.coin_quota = 0;
        s.coins_collected = 0;
        int curr_x = rand_randn(rg, h.main_width - 4) + 2;
        int curr_y = 0;
        int margin_x = 3;
        float enemy_prob = h.options.distribution_mode == EasyMode ? .2 : .5;
        for (int i = 0; i < num_platforms; i++) {
            int delta_y = choose_delta_y(c);
            bool can_spawn_enemy = (curr_x >= margin_x) && (curr_x <= h.main_width - margin_x);
            if (can_spawn_enemy && (rand_rand01(rg) < enemy_prob)) {
                // two RNG draws inside one argument list: g++ evaluates the arguments right to left,
                // so the velocity sign is drawn before the height offset (checked against the oracle)
                float evx = (float)(.15 * (rand_randn(rg, 2) * 2 - 1));
                float ey = (float)(curr_y + rand_randn(rg, 2) + 2 + .5);
                int ei = E::add_entity(c, (float)(curr_x + .5), ey, evx, 0, .5, ENEMY);
                Entity &ent = c.ents[ei];
                ent.image_type = ENEMY1;
                ent.smart_step = 1;
                ent.climber_spawn_x = (float)(curr_x + .5);
                E::match_aspect_ratio(c, ent);
            }
            curr_y += delta_y;
            int plat_len = 2 + rand_randn(rg, 10);
            int vx = rand_randn(rg, 2) * 2 - 1;
            if (curr_x < margin_x)
                vx = 1;
            if (curr_x > h.main_width - margin_x)
                vx = -1;
            int candidates[16];
            int ncand = 0;
            for (int j = 0; j < plat_len; j++) {
                int nx = curr_x + (j + 1) * vx;
                if (nx <= 0 || nx >= h.main_width - 1)
                    break;
                candidates[ncand++] = nx;
                E::set_obj(c, nx, curr_y, WALL_TOP);
            }
            if (ncand == 0) {
                h.err |= ERR_FASSERT;  // reference: choose_one on an empty vector
                return;
            }
            if (rand_rand01(rg) < .5 || i == num_platforms - 1) {
                int coin_x = candidates[rand_randn(rg, ncand)];
                E::add_entity(c, (float)(coin_x + .5), (float)(curr_y + 1.5), 0, 0, 0.3f, COIN);
                s.coin_quota += 1;
            }
            curr_x = candidates[rand_randn(rg, ncand)];
        }
    }
    // climber.cpp:233-236
    static PG_HD void choose_world_dim(Ctx &c) {
        c.h->main_width = c.h->options.distribution_mode == EasyMode ? 16 : 20;
        c.h->main_height = 64;
    }
    // climber.cpp:238-258
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        ClimberState &s = st(c);
        s.gravity = 0.2f;
        h.max_jump = 1.5;
        s.air_control = 0.15f;
        h.maxspeed = .5;
        s.has_support = 0;
        s.facing_right = 1;
        Entity &a = agent_of(c);
        a.rx = .5;
        a.ry = .5;
        a.x = 1 + a.rx;
        a.y = 1 + a.ry;
        E::choose_random_theme(c, a);
        s.wall_theme = rand_randn(*c.rng, NUM_WALL_THEMES);
        init_floor_and_walls(c);
        generate_platforms(c);
    }
    static PG_HD bool can_support(Ctx &c, int obj) { return is_wall(obj) || obj == c.oob; }
    // climber.cpp:264-268 (also pins visibility to the world width)
    static PG_HD void choose_center(Ctx &c, float &cx, float &cy) {
        cx = (float)(c.h->main_width / 2.0);
        cy = (float)((double)agent_of(c).y + c.h->main_width / 2.0 - (double)(5 * agent_of(c).ry));
        c.h->visibility = (float)c.h->main_width;
    }
    // climber.cpp:270-291
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        EnvHdr &h = *c.h;
        ClimberState &s = st(c);
        Entity &a = agent_of(c);
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
        s.has_support = can_support(c, obj_below_1) || can_support(c, obj_below_2);
        if (s.has_support && h.action_vy == 1)
            h.action_vy = 1;
        else
            h.action_vy = 0;
    }
    // climber.cpp:293-320
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
                if (ent.x > ent.climber_spawn_x + PATROL_RANGE)
                    ent.vx = (float)(-1 * pg_dfabs((double)ent.vx));
                else if (ent.x < ent.climber_spawn_x - PATROL_RANGE)
                    ent.vx = (float)pg_dfabs((double)ent.vx);
                ent.image_type = h.cur_time / 5 % 2 == 0 ? ENEMY1 : ENEMY2;
                ent.is_reflected = ent.vx < 0;
            }
        }
        if (st(c).coin_quota == st(c).coins_collected) {
            h.done = 1;
            h.reward += COMPLETION_BONUS;
            h.level_complete = 1;
        }
    }
};

}  // namespace pg
