// Leaper on the device engine. Behaviour restated from games/leaper.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct LeaperState {
    int32_t bottom_road_y;
    int32_t n_road;
    float road_lane_speeds[8];
    int32_t bottom_water_y;
    int32_t n_water;
    float water_lane_speeds[8];
    int32_t goal_y;
};

struct LeaperGame : Defaults<LeaperGame>, DrawDefaults<LeaperGame> {
    using E = Engine<LeaperGame>;
    static constexpr int ENT_CAP = 160;
    static constexpr int GRID_CAP = 20 * 20;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 160;
    static constexpr int MAX_ROT_BLITS = 4;
    static constexpr int MAX_VIEW_CELLS = 20;
    static constexpr const char *NAME = "leaper";
    // is_blocked / is_blocked_ents / will_reflect are the engine defaults here: only an entity typed WALL_OBJ or as the out-of-bounds object could block
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    // leaper.cpp:6-22
    static constexpr int LOG = 1, ROAD = 2, WATER = 3, CAR = 4, FINISH_LINE = 5;
    static constexpr float MONSTER_RADIUS = 0.25;
    static constexpr float LOG_RADIUS = 0.45f;
    static constexpr int GOAL_REWARD = 10;
    static constexpr int NSTEP = 5;
    static constexpr int FROG_ANIMATION_FRAMES = NSTEP;
    static constexpr float MAX_SPEED = (float)(2 / (NSTEP - 1.0));
    static constexpr float VEL_DECAY = MAX_SPEED / NSTEP;

    static PG_HD LeaperState &st(Ctx &c) { return game_state<LeaperState>(c); }

    // leaper.cpp:36-40
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->maxspeed = MAX_SPEED;
        c.h->timeout = 500;
    }
    // leaper.cpp:71-77
    static PG_HD float get_tile_aspect_ratio(Ctx &c, int ei) { return c.ents[ei].type == FINISH_LINE ? 1.0f : 0.0f; }
    // leaper.cpp:79-87 (does not call the base)
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        const Entity &obj = c.ents[oi];
        const Entity &a = agent_of(c);
        if (obj.type == CAR) {
            c.h->done = 1;
        } else if (obj.type == FINISH_LINE && a.vx == 0 && a.vy == 0) {
            c.h->reward += GOAL_REWARD;
            c.h->done = 1;
            c.h->level_complete = 1;
        }
    }
    static PG_HD bool should_preserve_type_themes(Ctx &c, int type) { return type == PLAYER; }
    // leaper.cpp:97-103
    static PG_HD float rand_sign(Ctx &c) { return rand_rand01(*c.rng) < 0.5 ? 1.0f : -1.0f; }
    // leaper.cpp:105-118
    static PG_HD void choose_world_dim(Ctx &c) {
        int dist_diff = c.h->options.distribution_mode;
        int world_dim = 20;
        if (dist_diff == EasyMode)
            world_dim = 9;
        else if (dist_diff == HardMode)
            world_dim = 15;
        c.h->main_width = world_dim;
        c.h->main_height = world_dim;
    }
    static PG_HD int choose_extra_space(Ctx &c) { return c.h->options.distribution_mode == EasyMode ? 0 : rand_randn(*c.rng, 2); }

    // leaper.cpp:184-212
    static PG_HD void spawn_entities(Ctx &c) {
        EnvHdr &h = *c.h;
        LeaperState &s = st(c);
        for (int lane = 0; lane < s.n_road; lane++) {
            float speed = s.road_lane_speeds[lane];
            float spawn_prob = (float)(pg_dfabs((double)speed) / 6.0);
            if (rand_rand01(*c.rng) < spawn_prob) {
                float x = speed > 0 ? (-1 * MONSTER_RADIUS) : (h.main_width + MONSTER_RADIUS);
                if (h.n_ents >= c.ent_cap) {
                    h.err |= ERR_ENTITY_OVERFLOW;
                    continue;
                }
                Entity &m = c.ents[h.n_ents];  // built past the end of the list, appended only if it fits
                entity_init(m, x, (float)(s.bottom_road_y + lane + 0.5), speed, 0, 2 * MONSTER_RADIUS, MONSTER_RADIUS, CAR);
                E::choose_random_theme(c, m);
                if (speed < 0)
                    m.rotation = PI_F;
                if (!E::has_any_collision(c, m))
                    E::push_entity(c);
            }
        }
        for (int lane = 0; lane < s.n_water; lane++) {
            float speed = s.water_lane_speeds[lane];
            float spawn_prob = (float)(pg_dfabs((double)speed) / 2.0);
            if (rand_rand01(*c.rng) < spawn_prob) {
                float x = speed > 0 ? (-1 * LOG_RADIUS) : (h.main_width + LOG_RADIUS);
                if (h.n_ents >= c.ent_cap) {
                    h.err |= ERR_ENTITY_OVERFLOW;
                    continue;
                }
                Entity &m = c.ents[h.n_ents];
                entity_init(m, x, (float)(s.bottom_water_y + lane + 0.5), speed, 0, LOG_RADIUS, LOG_RADIUS, LOG);
                if (!E::has_any_collision(c, m))
                    E::push_entity(c);
            }
        }
    }

    // leaper.cpp:124-182
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        LeaperState &s = st(c);
        h.options.center_agent = 0;
        agent_of(c).y = agent_of(c).ry;
        float min_car_speed = 0.05f, max_car_speed = 0.2f, min_log_speed = 0.05f, max_log_speed = 0.1f;
        if (h.options.distribution_mode == EasyMode) {
            min_car_speed = 0.03f;
            max_car_speed = 0.12f;
            min_log_speed = 0.025f;
            max_log_speed = 0.075f;
        } else if (h.options.distribution_mode == ExtremeMode) {
            min_car_speed = 0.1f;
            max_car_speed = 0.3f;
            min_log_speed = 0.1f;
            max_log_speed = 0.2f;
        }
        s.bottom_road_y = choose_extra_space(c) + 1;
        int max_diff = h.options.distribution_mode == EasyMode ? 3 : 4;
        int difficulty = rand_randn(*c.rng, max_diff + 1);
        int extra_lane_option = h.options.distribution_mode == EasyMode ? 0 : rand_randn(*c.rng, 4);
        int num_road_lanes = difficulty + (extra_lane_option == 2 ? 1 : 0);
        s.n_road = 0;
        for (int lane = 0; lane < num_road_lanes; lane++) {
            // operand order as compiled by g++ for `rand_sign() * rand_gen.randrange(..)`: left first
            float sg = rand_sign(c);
            float mag = rand_randrange(*c.rng, min_car_speed, max_car_speed);
            s.road_lane_speeds[s.n_road++] = sg * mag;
            E::fill_elem(c, 0, s.bottom_road_y + lane, h.main_width, 1, ROAD);
        }
        s.bottom_water_y = s.bottom_road_y + num_road_lanes + choose_extra_space(c) + 1;
        s.n_water = 0;
        int num_water_lanes = difficulty + (extra_lane_option == 3 ? 1 : 0);
        int curr_sign = (int)rand_sign(c);
        for (int lane = 0; lane < num_water_lanes; lane++) {
            s.water_lane_speeds[s.n_water++] = curr_sign * rand_randrange(*c.rng, min_log_speed, max_log_speed);
            curr_sign *= -1;
            E::fill_elem(c, 0, s.bottom_water_y + lane, h.main_width, 1, WATER);
        }
        s.goal_y = s.bottom_water_y + num_water_lanes + 1;
        const float lim = h.main_width / (min_car_speed < min_log_speed ? min_car_speed : min_log_speed);
        PG_PHASE_RESET(c);
        {
            PG_PHASE_BEGIN(c);
            for (int i = 0; (float)i < lim; i++) {
                spawn_entities(c);
                PG_PHASE_END(c, 0);
                E::step_entities(c);
                PG_PHASE_END(c, 1);
            }
            PG_PHASE_NOTE(c, 2, lim);
            PG_PHASE_NOTE(c, 3, c.h->n_ents);
        }
        E::add_entity_rxy(c, (float)(h.main_width / 2.0), (float)(s.goal_y - .5), 0, 0, (float)(h.main_width / 2.0), .5, FINISH_LINE);
    }

    // leaper.cpp:214-220 — sign() resolves to the double overload (cpp-utils.h:43)
    static PG_HD void decay_vel(float &vel) {
        float vel_sign = (float)pg_sign(1.0 * vel);
        vel = (float)(pg_dfabs((double)vel) - (double)VEL_DECAY);
        if (vel < 0)
            vel = 0;
        vel = vel * vel_sign;
    }
    // leaper.cpp:222-237
    static PG_HD void update_agent_velocity(Ctx &c) {
        EnvHdr &h = *c.h;
        Entity &a = agent_of(c);
        if (a.vx == 0 && a.vy == 0) {
            if (h.action_vx != 0) {
                a.vx = h.maxspeed * h.action_vx;
                a.image_theme = 1;
                a.rotation = (a.vx > 0 ? 1 : -1) * PI_F / 2;
            } else if (h.action_vy != 0) {
                a.vy = h.maxspeed * h.action_vy;
                a.image_theme = 1;
                a.rotation = a.vy > 0 ? 0 : PI_F;
            }
        }
        decay_vel(a.vx);
        decay_vel(a.vy);
    }
    // leaper.cpp:239-245
    static PG_HD bool get_adjusted_image_rect(Ctx &c, int type, double *adj) {
        if (type == PLAYER) {
            adj[0] = 0;
            adj[1] = -.275;
            adj[2] = 1;
            adj[3] = 1.55;
            return true;
        }
        return false;
    }
    // leaper.cpp:247-282
    static PG_HD void game_step(Ctx &c) {
        EnvHdr &h = *c.h;
        if (agent_of(c).image_theme >= 1)
            agent_of(c).image_theme = (agent_of(c).image_theme + 1) % FROG_ANIMATION_FRAMES;
        E::basic_game_step(c);
        spawn_entities(c);
        bool standing_on_log = false;
        float log_vx = 0.0;
        Entity &a = agent_of(c);
        float margin = -1 * a.rx;
        {
            // the reference's ascending loop leaves the LAST overlapping log in log_vx: find it warp-wide
            const Entity *ents = c.ents;
            const Entity av = a;
            const int li = pg_scan_down(h.n_ents, [=](int i) { return ents[i].type == LOG && E::has_collision(av, ents[i], margin); });
            if (li >= 0) {
                standing_on_log = true;
                log_vx = c.ents[li].vx;
            }
        }
        if (E::get_obj(c, (int)a.x, (int)a.y) == WATER) {
            if (!standing_on_log && a.vx == 0 && a.vy == 0)
                h.done = 1;
        }
        if (standing_on_log)
            a.x += log_vx;
        if (E::is_out_of_bounds(c, a))
            h.done = 1;
    }
};

}  // namespace pg
