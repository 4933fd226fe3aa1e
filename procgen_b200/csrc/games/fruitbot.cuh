// FruitBot on the device engine. Behaviour restated from games/fruitbot.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct FruitBotState {
    float min_dim, bullet_vscale;
    int32_t last_fire_time;
};

struct FruitBotGame : Defaults<FruitBotGame>, DrawDefaults<FruitBotGame> {
    using E = Engine<FruitBotGame>;
    static constexpr int ENT_CAP = 160;
    static constexpr int GRID_CAP = 20 * 60;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 256;  // barriers are tiled; on-screen tiles only (measured peak 156)
    static constexpr int MAX_ROT_BLITS = 2;
    static constexpr int MAX_VIEW_CELLS = 24;
    static constexpr const char *NAME = "fruitbot";
    // superset of the types is_blocked and will_reflect accept (barriers are entities)
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob || t == OUT_OF_BOUNDS_WALL || t == BARRIER; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    // fruitbot.cpp:8-23
    static constexpr float COMPLETION_BONUS = 10.0;
    static constexpr int POSITIVE_REWARD = 1, PENALTY = -4;
    static constexpr int BARRIER = 1, OUT_OF_BOUNDS_WALL = 2, PLAYER_BULLET = 3, BAD_OBJ = 4, GOOD_OBJ = 7, LOCKED_DOOR = 10, LOCK = 11, PRESENT = 12;
    static constexpr int KEY_DURATION = 8;
    static constexpr float DOOR_ASPECT_RATIO = 3.25;

    static PG_HD FruitBotState &st(Ctx &c) { return game_state<FruitBotState>(c); }

    // fruitbot.cpp:31-41
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->mixrate = .5;
        c.h->maxspeed = 0.85f;
        st(c).min_dim = 5;
        st(c).bullet_vscale = .5;
        c.h->bg_tile_ratio = -1;
        c.h->out_of_bounds_object = OUT_OF_BOUNDS_WALL;
    }
    static PG_HD bool will_reflect(Ctx &c, int src, int target) { return (src == BAD_OBJ && (target == BARRIER || target == WALL_OBJ)); }
    static PG_HD bool is_blocked(Ctx &c, int src, int target, bool is_horizontal) {
        return Defaults<FruitBotGame>::is_blocked(c, src, target, is_horizontal) || (c.ents[src].type == PLAYER && target == OUT_OF_BOUNDS_WALL);
    }
    // fruitbot.cpp:87-94
    static PG_HD float get_tile_aspect_ratio(Ctx &c, int ei) {
        int t = c.ents[ei].type;
        if (t == BARRIER)
            return 1;
        if (t == LOCKED_DOOR)
            return DOOR_ASPECT_RATIO;
        return 0;
    }
    // fruitbot.cpp:96-117
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        Entity &obj = c.ents[oi];
        EnvHdr &h = *c.h;
        if (obj.type == BARRIER) {
            h.done = 1;
        } else if (obj.type == BAD_OBJ) {
            h.reward += PENALTY;
            obj.will_erase = 1;
        } else if (obj.type == LOCKED_DOOR) {
            h.done = 1;
        } else if (obj.type == GOOD_OBJ) {
            h.reward += POSITIVE_REWARD;
            obj.will_erase = 1;
        } else if (obj.type == PRESENT) {
            h.reward += COMPLETION_BONUS;
            h.done = 1;
            h.level_complete = 1;
        }
    }
    // fruitbot.cpp:119-137
    static PG_HD void handle_collision(Ctx &c, int si, int ti) {
        Entity &src = c.ents[si];
        Entity &target = c.ents[ti];
        if (src.type == PLAYER_BULLET) {
            if (target.type == BARRIER) {
                src.will_erase = 1;
            } else if (target.type == LOCK) {
                src.will_erase = 1;
                target.will_erase = 1;
                for (int i = 0; i < c.h->n_ents; i++) {
                    Entity &ent = c.ents[i];
                    if (ent.type == LOCKED_DOOR && pg_dfabs((double)(ent.y - target.y)) < 1) {
                        ent.will_erase = 1;
                        break;
                    }
                }
            }
        }
    }
    // fruitbot.cpp:143-147
    static PG_HD void choose_center(Ctx &c, float &cx, float &cy) {
        cx = (float)(c.h->main_width / 2.0);
        cy = (float)((double)agent_of(c).y + c.h->main_width / 2.0 - (double)(2 * agent_of(c).ry));
        c.h->visibility = (float)c.h->main_width;
    }
    // fruitbot.cpp:149-157
    static PG_HD void choose_world_dim(Ctx &c) {
        c.h->main_width = c.h->options.distribution_mode == EasyMode ? 10 : 20;
        c.h->main_height = 60;
    }
    // fruitbot.cpp:159-163
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        c.h->action_vx = move_action / 3 - 1;
        c.h->action_vy = 0.2f;
        c.h->action_vrot = 0;
    }
    // fruitbot.cpp:165-199
    static PG_HD void add_walls(Ctx &c, float ry, bool use_door, float min_pct) {
        EnvHdr &h = *c.h;
        MT19937 &rg = *c.rng;
        float rw = (float)h.main_width;
        float wall_ry = 0.3f;
        float lock_rx = .25;
        float lock_ry = 0.45f;
        float pct = (float)((double)min_pct + .2 * (double)rand_rand01(rg));
        if (use_door) {
            pct += 0.1f;
            float lock_pct_w = 2 * lock_rx / h.main_width;
            float door_pct_w = (wall_ry * 2 * DOOR_ASPECT_RATIO) / h.main_width;
            int num_doors = (int)pg_dceil((double)((pct - 2 * lock_pct_w) / door_pct_w));
            pct = 2 * lock_pct_w + door_pct_w * num_doors;
        }
        float gapw = pct * rw;
        float w1 = rand_rand01(rg) * (rw - gapw);
        float w2 = rw - w1 - gapw;
        E::add_entity_rxy(c, w1 / 2, ry, 0, 0, w1 / 2, wall_ry, BARRIER);
        E::add_entity_rxy(c, rw - w2 / 2, ry, 0, 0, w2 / 2, wall_ry, BARRIER);
        if (use_door) {
            int is_on_right = rand_randn(rg, 2);
            float lock_x = w1 + lock_rx + is_on_right * (gapw - 2 * lock_rx);
            float door_x = w1 + gapw / 2 - (is_on_right * 2 - 1) * lock_rx;
            E::add_entity_rxy(c, door_x, ry, 0, 0, gapw / 2 - lock_rx, wall_ry, LOCKED_DOOR);
            E::add_entity_rxy(c, lock_x, ry - lock_ry + wall_ry, 0, 0, lock_rx, lock_ry, LOCK);
        }
    }
    // fruitbot.cpp:201-253
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        MT19937 &rg = *c.rng;
        st(c).last_fire_time = 0;
        int min_sep = 4;
        int num_walls = 10;
        int object_group_size = 6;
        int buf_h = 4;
        float door_prob = .125;
        float min_pct = .1;
        if (h.options.distribution_mode == EasyMode) {
            num_walls = 5;
            object_group_size = 2;
            door_prob = 0;
            min_pct = .2;
        }
        // RandGen::partition, randgen.cpp:33-41
        int partition[10];
        for (int i = 0; i < num_walls; i++) partition[i] = 0;
        const int px = h.main_height - min_sep * num_walls - buf_h;
        for (int i = 0; i < px; i++) partition[rand_randn(rg, num_walls)] += 1;
        int curr_h = 0;
        for (int q = 0; q < num_walls; q++) {
            int dy = min_sep + partition[q];
            curr_h += dy;
            bool use_door = (dy > 5) && rand_rand01(rg) < door_prob;
            add_walls(c, (float)curr_h, use_door, min_pct);
        }
        agent_of(c).y = agent_of(c).ry;
        int num_good = rand_randn(rg, 10) + 10;
        int num_bad = rand_randn(rg, 10) + 10;
        for (int i = 0; i < h.main_width; i++) {
            int pi = E::add_entity_rxy(c, (float)(i + .5), (float)(h.main_height - .5), 0, 0, .5, .5, PRESENT);
            E::choose_random_theme(c, c.ents[pi]);
        }
        E::spawn_entities(c, num_good, .5, GOOD_OBJ, 0, 0, (float)h.main_width, (float)h.main_height);
        E::spawn_entities(c, num_bad, .5, BAD_OBJ, 0, 0, (float)h.main_width, (float)h.main_height);
        for (int i = 0; i < h.n_ents; i++) {
            Entity &ent = c.ents[i];
            if (ent.type == GOOD_OBJ || ent.type == BAD_OBJ) {
                ent.image_theme = rand_randn(rg, object_group_size);
                E::fit_aspect_ratio(c, ent);
            }
        }
        agent_of(c).rotation = -1 * PI_F / 2;
    }
    // fruitbot.cpp:255-266
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        if (h.special_action == 1 && (h.cur_time - st(c).last_fire_time) >= KEY_DURATION) {
            float vx = 0;
            float vy = 1;
            Entity &a = agent_of(c);
            int bi = E::add_entity(c, a.x, a.y, vx * st(c).bullet_vscale, vy * st(c).bullet_vscale, .25, PLAYER_BULLET);
            c.ents[bi].expire_time = KEY_DURATION;
            c.ents[bi].collides_with_entities = 1;
            st(c).last_fire_time = h.cur_time;
        }
    }
};

}  // namespace pg
