// Plunder on the device engine. Behaviour restated from games/plunder.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct PlunderState {
    int32_t last_fire_time;
    int32_t lane_directions[5];
    int32_t target_bools[6];
    int32_t image_permutation[6];
    float lane_vels[5];
    int32_t num_lanes, num_current_ship_types, targets_hit, target_quota;
    float juice_left, r_scale, spawn_prob, legend_r, min_agent_x;
};

struct PlunderGame : Defaults<PlunderGame>, DrawDefaults<PlunderGame> {
    using E = Engine<PlunderGame>;
    static constexpr int ENT_CAP = 64;
    static constexpr int GRID_CAP = 20 * 20;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 64;
    static constexpr int MAX_ROT_BLITS = 4;
    static constexpr int MAX_VIEW_CELLS = 20;
    static constexpr const char *NAME = "plunder";
    static constexpr bool DRAWS_GRID = false;  // entities only; the grid stays all SPACE
    // is_blocked / is_blocked_ents / will_reflect are the engine defaults here: only an entity typed WALL_OBJ or as the out-of-bounds object could block
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    // plunder.cpp:8-15
    static constexpr float COMPLETION_BONUS = 10.0f;
    static constexpr float POSITIVE_REWARD = 1.0f;
    static constexpr int PLAYER_BULLET = 1, TARGET_LEGEND = 2, TARGET_BACKGROUND = 3, PANEL = 6, SHIP = 7;

    static PG_HD PlunderState &st(Ctx &c) { return game_state<PlunderState>(c); }

    // plunder.cpp:34-44
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->timeout = 4000;
        c.h->main_width = 20;
        c.h->main_height = 20;
        c.h->mixrate = .5;
        c.h->maxspeed = 0.85f;
        c.h->has_useful_vel_info = 0;
    }
    // plunder.cpp:66-77: two HUD bars drawn after everything else
    template <class Frame>
    static PG_HD void make_overlay_blits(Ctx &c, Frame &f) {
        EnvHdr &h = *c.h;
        PlunderState &s = st(c);
        double r[4];
        Raster<PlunderGame, Frame>::abs_rect(f.cam, .25, .25, h.main_width * s.juice_left, .5, r);
        make_solid_blit(f.overlay[f.n_overlay++], r[0], r[1], r[2], r[3], (66u << 16) | (245u << 8) | 135u);
        Raster<PlunderGame, Frame>::abs_rect(f.cam, .25, .75, (float)(h.main_width * (s.targets_hit * 1.0 / s.target_quota)), .5, r);
        make_solid_blit(f.overlay[f.n_overlay++], r[0], r[1], r[2], r[3], (245u << 16) | (66u << 8) | 144u);
    }
    static PG_HD bool should_preserve_type_themes(Ctx &c, int type) { return type == SHIP; }
    // plunder.cpp:87-110
    static PG_HD void handle_collision(Ctx &c, int si, int ti) {
        Entity &src = c.ents[si];
        Entity &target = c.ents[ti];
        PlunderState &s = st(c);
        if (src.type == PLAYER_BULLET) {
            if (target.type == SHIP) {
                target.will_erase = 1;
                src.will_erase = 1;
                if (s.target_bools[target.image_theme]) {
                    s.targets_hit += 1;
                    c.h->reward += POSITIVE_REWARD;
                    s.juice_left += 0.1f;
                } else {
                    s.juice_left -= 0.1f;
                }
            } else if (target.type == PANEL) {
                src.will_erase = 1;
            }
            if (target.will_erase) {
                float tx = target.x, ty = target.y, tvx = target.vx / 2, tvy = target.vy / 2, tr = (float)(.5 * target.rx);
                E::add_entity(c, tx, ty, tvx, tvy, tr, EXPLOSION);
            }
        }
    }
    // plunder.cpp:112-116
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        c.h->action_vx = move_action / 3 - 1;
        c.h->action_vy = 0;
        c.h->action_vrot = 0;
    }
    // plunder.cpp:118-192
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        PlunderState &s = st(c);
        MT19937 &rg = *c.rng;
        agent_of(c).image_type = SHIP;
        s.juice_left = 1;
        s.targets_hit = 0;
        s.target_quota = 20;
        s.spawn_prob = 0.06f;
        s.r_scale = h.options.distribution_mode == EasyMode ? 1.5f : 1.0f;
        const int num_total_ship_types = 6;
        s.num_lanes = 5;
        // RandGen::choose_n(image_idxs, 6), randgen.cpp:53-70
        {
            int rem[6], nrem = 6;
            for (int i = 0; i < 6; i++) rem[i] = i;
            for (int k = 0; k < num_total_ship_types; k++) {
                int idx = rand_randn(rg, nrem);
                s.image_permutation[k] = rem[idx];
                for (int j = idx; j < nrem - 1; j++) rem[j] = rem[j + 1];
                nrem--;
            }
        }
        s.num_current_ship_types = 2;
        for (int i = 0; i < num_total_ship_types; i++) s.target_bools[i] = 0;
        for (int i = 0; i < s.num_current_ship_types / 2; i++) s.target_bools[s.image_permutation[i]] = 1;
        for (int i = 0; i < s.num_lanes; i++) {
            s.lane_directions[i] = rand_rand01(rg) < .5;
            s.lane_vels[i] = (float)(.15 + .1 * (double)rand_rand01(rg));
        }
        int num_panels = h.options.distribution_mode == EasyMode ? 0 : rand_randn(rg, 4);
        float panel_width = 1.2f;
        if (panel_width > 0) {
            for (int i = 0; i < num_panels; i++)
                E::spawn_entity_rxy(c, panel_width, .5, PANEL, 0, (float)(.25 * h.main_height), (float)h.main_width, (float)(.25 * h.main_height));
        }
        float key_scale = 1.5;
        s.legend_r = 2;
        E::add_entity(c, s.legend_r, s.legend_r, 0, 0, s.legend_r, TARGET_BACKGROUND);
        int ei = E::add_entity(c, s.legend_r, s.legend_r, 0, 0, s.r_scale * key_scale, TARGET_LEGEND);
        Entity &ent = c.ents[ei];
        ent.image_theme = s.image_permutation[0];
        ent.image_type = SHIP;
        E::match_aspect_ratio(c, ent);
        ent.rotation = PI_F / 2;
        s.last_fire_time = 0;
        h.options.center_agent = 0;
        Entity &a = agent_of(c);
        a.rx = s.r_scale;
        a.rotation = -1 * PI_F / 2;
        a.image_theme = s.image_permutation[rand_randn(rg, s.num_current_ship_types / 2) + s.num_current_ship_types / 2];
        E::match_aspect_ratio(c, a);
        E::reposition_agent(c);
        a.y = 1 + a.ry;
        s.min_agent_x = 2 * s.legend_r + a.rx;
        if (a.x < s.min_agent_x)
            a.x = s.min_agent_x;
    }
    // plunder.cpp:194-245
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        PlunderState &s = st(c);
        MT19937 &rg = *c.rng;
        s.juice_left -= 0.0015f;
        if (rand_rand01(rg) < s.spawn_prob) {
            float ent_r = s.r_scale;
            int lane = rand_randn(rg, s.num_lanes);
            float ent_y = (float)((lane * .11 + .4) * (double)(h.main_height / 2 - ent_r) + h.main_height / 2);
            float moves_right = (float)s.lane_directions[lane];
            float ent_vx = s.lane_vels[lane] * (moves_right ? 1 : -1);
            if (h.n_ents >= c.ent_cap) {
                h.err |= ERR_ENTITY_OVERFLOW;
            } else {
                Entity &ent = c.ents[h.n_ents];
                entity_init(ent, 0, ent_y, ent_vx, 0, ent_r, ent_r, SHIP);
                ent.image_type = SHIP;
                ent.image_theme = s.image_permutation[rand_randn(rg, s.num_current_ship_types)];
                E::match_aspect_ratio(c, ent);
                ent.x = moves_right ? -1 * ent_r : (h.main_width + ent_r);
                ent.is_reflected = !moves_right;
                if (!E::has_any_collision(c, ent))
                    E::push_entity(c);
            }
        }
        if (h.special_action == 1 && (h.cur_time - s.last_fire_time) >= 3) {
            Entity &a = agent_of(c);
            int bi = E::add_entity(c, a.x, a.y, 0, 1, .25, PLAYER_BULLET);
            c.ents[bi].collides_with_entities = 1;
            c.ents[bi].expire_time = 50;
            s.last_fire_time = h.cur_time;
            s.juice_left -= 0.02f;
        }
        if (s.juice_left <= 0) {
            h.done = 1;
        } else if (s.juice_left >= 1) {
            s.juice_left = 1;
        }
        if (s.targets_hit >= s.target_quota) {
            h.done = 1;
            h.reward += COMPLETION_BONUS;
            h.level_complete = 1;
        }
        if (agent_of(c).x < s.min_agent_x)
            agent_of(c).x = s.min_agent_x;
    }
};

}  // namespace pg
