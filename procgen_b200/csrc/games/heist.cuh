// Heist on the device engine. Behaviour restated from games/heist.cpp (cited per function).
#pragma once
#include "../pg_mazegen.cuh"
#include "../pg_raster.cuh"

namespace pg {

struct HeistState {
    int32_t world_dim;
    int32_t num_keys;
    int32_t has_keys[4];
};

struct HeistGame : Defaults<HeistGame>, DrawDefaults<HeistGame> {
    using E = Engine<HeistGame>;
    static constexpr int ENT_CAP = 32;
    static constexpr int GRID_CAP = 23 * 23;
    static constexpr int SCRATCH_WORDS = 8192;   // MazeGen::words_needed(23) = 6717
    static constexpr int MAX_VISIBLE_ENTS = 64;
    static constexpr int MAX_ROT_BLITS = 8;
    static constexpr int MAX_VIEW_CELLS = 13;    // hard: whole 13x13 world; memory mode is centred (11)
    static constexpr const char *NAME = "heist";
    // locked doors are entities and block without their key (is_blocked_ents); otherwise the defaults
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob || t == LOCKED_DOOR; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    // heist.cpp:10-15
    static constexpr float COMPLETION_BONUS = 10.0f;
    static constexpr int LOCKED_DOOR = 1, KEY = 2, EXIT = 9, KEY_ON_RING = 11;

    static PG_HD HeistState &st(Ctx &c) { return game_state<HeistState>(c); }

    // heist.cpp:24-34
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->has_useful_vel_info = 0;
        c.h->main_width = 20;
        c.h->main_height = 20;
        c.h->out_of_bounds_object = WALL_OBJ;
        c.h->visibility = 8.0;
    }
    // heist.cpp:40-42
    static PG_HD bool should_preserve_type_themes(Ctx &c, int type) { return type == KEY || type == LOCKED_DOOR; }
    // heist.cpp:66-71
    static PG_HD bool is_blocked_ents(Ctx &c, int src, int target, bool is_horizontal) {
        const Entity &t = c.ents[target];
        if (t.type == LOCKED_DOOR)
            return !st(c).has_keys[t.image_theme];
        return Defaults<HeistGame>::is_blocked_ents(c, src, target, is_horizontal);
    }
    // heist.cpp:73-78
    static PG_HD bool should_draw_entity(Ctx &c, int ei) {
        const Entity &e = c.ents[ei];
        if (e.type == KEY_ON_RING)
            return st(c).has_keys[e.image_theme] != 0;
        return true;
    }
    // heist.cpp:80-97
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        Entity &obj = c.ents[oi];
        if (obj.type == EXIT) {
            c.h->done = 1;
            c.h->reward = COMPLETION_BONUS;
            c.h->level_complete = 1;
        } else if (obj.type == KEY) {
            obj.will_erase = 1;
            st(c).has_keys[obj.image_theme] = 1;
        } else if (obj.type == LOCKED_DOOR) {
            int door_num = obj.image_theme;
            if (st(c).has_keys[door_num])
                obj.will_erase = 1;
        }
    }
    // heist.cpp:99-116
    static PG_HD void choose_world_dim(Ctx &c) {
        int dist_diff = c.h->options.distribution_mode;
        if (dist_diff == EasyMode)
            st(c).world_dim = 9;
        else if (dist_diff == HardMode)
            st(c).world_dim = 13;
        else if (dist_diff == MemoryMode)
            st(c).world_dim = 23;
        c.h->maxspeed = .75;
        c.h->main_width = st(c).world_dim;
        c.h->main_height = st(c).world_dim;
    }
    // heist.cpp:118-204
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        HeistState &s = st(c);
        MT19937 &rg = *c.rng;
        const int world_dim = s.world_dim;
        int min_maze_dim = 5;
        int max_diff = (world_dim - min_maze_dim) / 2;
        int difficulty = rand_randn(rg, max_diff + 1);
        h.options.center_agent = h.options.distribution_mode == MemoryMode;
        if (h.options.distribution_mode == MemoryMode)
            s.num_keys = rand_randn(rg, 4);
        else
            s.num_keys = difficulty + rand_randn(rg, 2);
        if (s.num_keys > 3)
            s.num_keys = 3;
        for (int i = 0; i < 4; i++) s.has_keys[i] = 0;
        int maze_dim = difficulty * 2 + min_maze_dim;
        float maze_scale = (float)(h.main_height / (world_dim * 1.0));
        agent_of(c).rx = (float)(.375 * maze_scale);
        agent_of(c).ry = (float)(.375 * maze_scale);
        float r_ent = maze_scale / 2;
        MazeGen mg;
        mg.init(c, maze_dim);
        mg.generate_maze_with_doors(s.num_keys);
        agent_of(c).x = -1;
        agent_of(c).y = -1;
        int off_x = rand_randn(rg, world_dim - maze_dim + 1);
        int off_y = rand_randn(rg, world_dim - maze_dim + 1);
        {
            int16_t *g = c.grid;
            pg_warp_for(h.grid_size, [=](int i) { g[i] = (int16_t)WALL_OBJ; });
        }
        for (int i = 0; i < maze_dim; i++) {
            for (int j = 0; j < maze_dim; j++) {
                int x = off_x + i;
                int y = off_y + j;
                int obj = mg.grid_get(i + MAZE_OFFSET, j + MAZE_OFFSET);
                float obj_x = (float)((x + .5) * maze_scale);
                float obj_y = (float)((y + .5) * maze_scale);
                if (obj != WALL_OBJ)
                    E::set_obj(c, x, y, SPACE);
                if (obj >= KEY_OBJ) {
                    int ei = E::spawn_entity(c, (float)(.375 * maze_scale), KEY, maze_scale * x, maze_scale * y, maze_scale, maze_scale);
                    c.ents[ei].image_theme = obj - KEY_OBJ - 1;
                    E::match_aspect_ratio(c, c.ents[ei]);
                } else if (obj >= DOOR_OBJ) {
                    int ei = E::add_entity(c, obj_x, obj_y, 0, 0, r_ent, LOCKED_DOOR);
                    c.ents[ei].image_theme = obj - DOOR_OBJ - 1;
                } else if (obj == EXIT_OBJ) {
                    int ei = E::spawn_entity(c, (float)(.375 * maze_scale), EXIT, maze_scale * x, maze_scale * y, maze_scale, maze_scale);
                    E::match_aspect_ratio(c, c.ents[ei]);
                } else if (obj == AGENT_OBJ) {
                    agent_of(c).x = obj_x;
                    agent_of(c).y = obj_y;
                }
            }
        }
        float ring_key_r = 0.03f;
        for (int i = 0; i < s.num_keys; i++) {
            int ei = E::add_entity(c, (float)(1 - ring_key_r * (2 * i + 1.25)), (float)(ring_key_r * .75), 0, 0, ring_key_r, KEY_ON_RING);
            Entity &ent = c.ents[ei];
            ent.image_theme = i;
            ent.image_type = KEY;
            ent.rotation = PI_F / 2;
            ent.render_z = 1;
            ent.use_abs_coords = 1;
            E::match_aspect_ratio(c, ent);
        }
    }
    // heist.cpp:206-210
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        entity_face_direction(agent_of(c), c.h->action_vx, c.h->action_vy);
    }
};

}  // namespace pg
