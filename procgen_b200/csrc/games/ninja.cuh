// Ninja on the device engine. Behaviour restated from games/ninja.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct NinjaState {
    int32_t has_support, facing_right, last_fire_time, wall_theme;
    float gravity, air_control, jump_charge, jump_charge_inc;
};

struct Ninja : Defaults<Ninja>, DrawDefaults<Ninja> {
    using E = Engine<Ninja>;
    static constexpr int ENT_CAP = 64;
    static constexpr int GRID_CAP = 64 * 64;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 64;
    static constexpr int MAX_ROT_BLITS = 0;
    static constexpr int MAX_VIEW_CELLS = 20;  // visibility 16: int(c-9)..int(c+9)
    static constexpr int FULL_VIEW_CELLS = 64;  // center_agent = false: the whole world (basic-abstract-game.cpp:819-838)
    static constexpr const char *NAME = "ninja";

    // ninja.cpp:9-21
    static constexpr float GOAL_REWARD = 10.0f;
    static constexpr int GOAL = 1, BOMB = 6, THROWING_STAR = 7, PLAYER_JUMP = 9, PLAYER_RIGHT1 = 12, PLAYER_RIGHT2 = 13, FIRE = 14;
    static constexpr int WALL_MID = 20;
    static constexpr int NUM_WALL_THEMES = 3;

    static PG_HD NinjaState &st(Ctx &c) { return game_state<NinjaState>(c); }
    static PG_HD bool is_wall(int type) { return type == WALL_MID; }

    // ninja.cpp:34-40
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->main_width = 64;
        c.h->main_height = 64;
        c.h->out_of_bounds_object = WALL_MID;
    }
    // ninja.cpp:77-87
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        int t = c.ents[oi].type;
        if (t == EXPLOSION) {
            c.h->done = 1;
        } else if (t == GOAL) {
            c.h->reward += GOAL_REWARD;
            c.h->level_complete = 1;
            c.h->done = 1;
        }
    }
    // ninja.cpp:89-107
    static PG_HD void handle_grid_collision(Ctx &c, int oi, int type, int i, int j) {
        Entity &obj = c.ents[oi];
        if (obj.type == PLAYER) {
            if (type == FIRE)
                c.h->done = 1;
            else if (type == BOMB)
                c.h->done = 1;
        } else if (obj.type == THROWING_STAR) {
            if (type == BOMB) {
                obj.will_erase = 1;
                E::set_obj(c, i, j, SPACE);
                E::add_entity(c, (float)(i + .5), (float)(j + .5), 0, 0, .5, EXPLOSION);
            }
            if (is_wall(type))
                c.ents[oi].will_erase = 1;
        }
    }
    // ninja.cpp:109-123
    static PG_HD void update_agent_velocity(Ctx &c) {
        EnvHdr &h = *c.h;
        NinjaState &s = st(c);
        Entity &a = agent_of(c);
        float mixrate_x = s.has_support ? h.mixrate : (h.mixrate * s.air_control);
        a.vx = (1 - mixrate_x) * a.vx + mixrate_x * h.maxspeed * h.action_vx;
        if (h.action_vy < 1 && s.jump_charge > 0) {
            a.vy = s.jump_charge * h.max_jump;
            s.jump_charge = 0;
        }
        if (!s.has_support) {
            if (a.vy > -2)
                a.vy -= s.gravity;
        }
    }
    static PG_HD int theme_for_grid_obj(Ctx &c, int type) { return is_wall(type) ? st(c).wall_theme : 0; }
    // ninja.cpp:140-154 (has a side effect: stars stick to walls)
    static PG_HD bool is_blocked(Ctx &c, int src, int target, bool is_horizontal) {
        if (is_wall(target)) {
            Entity &s = c.ents[src];
            if (s.type == PLAYER) {
                return true;
            } else if (s.type == THROWING_STAR) {
                s.vx = 0;
                s.vy = 0;
                return true;
            }
        }
        return Defaults<Ninja>::is_blocked(c, src, target, is_horizontal);
    }
    // ninja.cpp:156-166
    static PG_HD int image_for_type(Ctx &c, int type) {
        if (type == PLAYER) {
            if (pg_dfabs((double)agent_of(c).vx) < .01 && c.h->action_vx == 0 && st(c).has_support)
                return PLAYER;
            return (c.h->cur_time / 5 % 2 == 0 || !st(c).has_support) ? PLAYER_RIGHT1 : PLAYER_RIGHT2;
        }
        return Defaults<Ninja>::image_for_type(c, type);
    }
    // ninja.cpp:168-177: jump-charge bar
    template <class Frame>
    static PG_HD void make_overlay_blits(Ctx &c, Frame &f) {
        float bar_height = 3 * st(c).jump_charge;
        double r[4];
        Raster<Ninja, Frame>::abs_rect(f.cam, .25, (float)((double)c.h->visibility - .5 - (double)bar_height), .5, bar_height, r);
        make_solid_blit(f.overlay[f.n_overlay++], r[0], r[1], r[2], r[3], (66u << 16) | (245u << 8) | 135u);
    }
    static PG_HD void fill_ground_block(Ctx &c, int x, int y, int dx, int dy) {
        if (dy <= 0)
            return;
        E::fill_elem(c, x, y, dx, dy - 1, WALL_MID);
        E::fill_elem(c, x, y + dy - 1, dx, 1, WALL_MID);
    }
    static PG_HD void init_floor_and_walls(Ctx &c) {
        int w = c.h->main_width, h = c.h->main_height;
        E::fill_elem(c, 0, 0, w, 1, WALL_MID);
        E::fill_elem(c, 0, 0, 1, h, WALL_MID);
        E::fill_elem(c, w - 1, 0, 1, h, WALL_MID);
        E::fill_elem(c, 0, h - 1, w, 1, WALL_MID);
    }
    // ninja.cpp:197-305
    static PG_HD void generate_coin_to_the_right(Ctx &c, int difficulty) {
        EnvHdr &h = *c.h;
        MT19937 &rg = *c.rng;
        int min_gap = difficulty - 1;
        int min_plat_w = 1;
        int inc_dy = 4;
        if (h.options.distribution_mode == EasyMode) {
            min_gap -= 1;
            if (min_gap < 0)
                min_gap = 0;
            min_plat_w = 3;
            inc_dy = 2;
        }
        float bomb_prob = (float)(.25 * (difficulty - 1));
        int max_gap_inc = difficulty == 1 ? 1 : 2;
        int num_sections = rand_randn(rg, difficulty) + difficulty;
        int start_x = 5;
        int curr_x = start_x;
        int curr_y = h.main_height / 2;
        int min_y = curr_y;
        int w = h.main_width;
        float _max_dy = h.max_jump * h.max_jump / (2 * st(c).gravity);
        int max_dy = (int)(_max_dy - .5);
        int prev_x, prev_y;
        fill_ground_block(c, 0, 0, start_x, curr_y);
        E::fill_elem(c, 0, curr_y + 8, start_x, h.main_height - curr_y - 8, WALL_MID);
        for (int i = 0; i < num_sections; i++) {
            prev_x = curr_x;
            prev_y = curr_y;
            int num_edges = rand_randn(rg, 2) + 1;
            int max_y = -1;
            int last_edge_y = -1;
            for (int j = 0; j < num_edges; j++) {
                curr_x = prev_x + j;
                if (curr_x + 15 >= w)
                    break;
                curr_y = prev_y;
                int dy = rand_randn(rg, inc_dy) + 1 + int(difficulty / 3);
                if (dy > max_dy)
                    dy = max_dy;
                if (curr_y >= h.main_height - 15) {
                    dy *= -1;
                } else if (curr_y >= 5 && rand_rand01(rg) < .4) {
                    dy *= -1;
                }
                curr_y += dy;
                if (curr_y < 3)
                    curr_y = 3;
                int diff = curr_y - last_edge_y;
                if ((diff < 0 ? -diff : diff) <= 1)
                    curr_y = last_edge_y + 2;
                int dx = min_plat_w + rand_randn(rg, 3);
                fill_ground_block(c, curr_x, curr_y - 1, dx, 1);
                curr_x += dx;
                curr_x += min_gap + rand_randn(rg, max_gap_inc + 1);
                if (curr_y > max_y)
                    max_y = curr_y;
                if (curr_y < min_y)
                    min_y = curr_y;
                last_edge_y = curr_y;
            }
            if (rand_rand01(rg) < bomb_prob)
                E::set_obj(c, rand_randn(rg, curr_x - prev_x + 1) + prev_x, max_y + 2, BOMB);
            int ceiling_height = 11;
            int ceiling_start = max_y - 1 + ceiling_height;
            fill_ground_block(c, prev_x, ceiling_start, curr_x - prev_x, h.main_height - ceiling_start);
        }
        int gi = E::add_entity(c, (float)(curr_x + .5), (float)(curr_y + .5), 0, 0, .5, GOAL);
        E::choose_random_theme(c, c.ents[gi]);
        fill_ground_block(c, curr_x, curr_y - 1, 1, 1);
        E::fill_elem(c, curr_x, curr_y + 6, 1, h.main_height - curr_y - 6, WALL_MID);
        int fire_y = min_y - 2;
        if (fire_y < 1)
            fire_y = 1;
        fill_ground_block(c, start_x, 0, h.main_width - start_x, fire_y);
        E::fill_elem(c, start_x, fire_y, h.main_width - start_x, 1, FIRE);
        E::fill_elem(c, curr_x + 1, 0, h.main_width - curr_x - 1, h.main_height, WALL_MID);
    }
    // ninja.cpp:307-341
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        NinjaState &s = st(c);
        s.gravity = 0.2f;
        h.max_jump = 1.5;
        s.air_control = 0.15f;
        h.maxspeed = .5;
        s.has_support = 0;
        s.facing_right = 1;
        s.jump_charge = 0;
        s.jump_charge_inc = .25;
        h.visibility = 16;
        Entity &a = agent_of(c);
        a.rx = .5;
        a.ry = .5;
        a.x = 1 + a.rx;
        a.y = h.main_height / 2 + a.ry;
        if (h.options.distribution_mode == EasyMode) {
            h.max_jump = 1.25;
            s.jump_charge_inc = 1;
            h.visibility = 10;
        }
        int max_difficulty = 3;
        int difficulty = rand_randn(*c.rng, max_difficulty) + 1;
        s.last_fire_time = 0;
        s.wall_theme = rand_randn(*c.rng, NUM_WALL_THEMES);
        init_floor_and_walls(c);
        generate_coin_to_the_right(c, difficulty);
    }
    static PG_HD bool can_support(Ctx &c, int obj) { return is_wall(obj) || obj == c.oob; }
    // ninja.cpp:347-378
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        EnvHdr &h = *c.h;
        NinjaState &s = st(c);
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
        if (s.has_support && h.action_vy == 1) {
            h.action_vy = 1;
            s.jump_charge += s.jump_charge_inc;
            if (s.jump_charge > 1)
                s.jump_charge = 1;
        } else {
            h.action_vy = 0;
        }
        if (!s.has_support)
            s.jump_charge = 0;
    }
    // ninja.cpp:380-413 — cos/sin are the double overloads
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        NinjaState &s = st(c);
        if (h.action_vx > 0)
            agent_of(c).is_reflected = 0;
        if (h.action_vx < 0)
            agent_of(c).is_reflected = 1;
        if (h.special_action > 0 && (h.cur_time - s.last_fire_time) >= 3) {
            float theta = 0;
            float bullet_vel = 1;
            if (h.special_action == 1)
                theta = 0;
            else if (h.special_action == 2)
                theta = PI_F / 4;
            else if (h.special_action == 3)
                theta = PI_F / 2;
            else if (h.special_action == 4)
                theta = -1 * PI_F / 4;
            if (agent_of(c).is_reflected)
                theta = PI_F - theta;
            Entity &a = agent_of(c);
            int bi = E::add_entity(c, a.x, a.y, (float)((double)bullet_vel * cos((double)theta)), (float)((double)bullet_vel * sin((double)theta)), .25, THROWING_STAR);
            c.ents[bi].collides_with_entities = 1;
            c.ents[bi].expire_time = 15;
            c.ents[bi].smart_step = 1;
            s.last_fire_time = h.cur_time;
        }
    }
};

}  // namespace pg
