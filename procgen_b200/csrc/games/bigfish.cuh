// BigFish on the device engine. Behaviour restated from games/bigfish.cpp (cited per function).
#pragma once
#include "../pg_raster.cuh"

namespace pg {

struct BigFishState {
    int32_t fish_eaten;
    float r_inc;
};

struct BigFish : Defaults<BigFish>, DrawDefaults<BigFish> {
    using E = Engine<BigFish>;
    static constexpr int ENT_CAP = 64;
    static constexpr int GRID_CAP = 20 * 20;
    static constexpr int SCRATCH_WORDS = 0;
    static constexpr int MAX_VISIBLE_ENTS = 64;
    static constexpr int MAX_ROT_BLITS = 0;
    static constexpr int MAX_VIEW_CELLS = 20;
    static constexpr const char *NAME = "bigfish";
    static constexpr bool DRAWS_GRID = false;  // entities only; the grid stays all SPACE
    // is_blocked / is_blocked_ents / will_reflect are the engine defaults here: only an entity typed WALL_OBJ or as the out-of-bounds object could block
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    // bigfish.cpp:8-16
    static constexpr int COMPLETION_BONUS = 10;
    static constexpr int POSITIVE_REWARD = 1;
    static constexpr int FISH = 2;
    static constexpr float FISH_MIN_R = .25;
    static constexpr float FISH_MAX_R = 2;
    static constexpr int FISH_QUOTA = 30;

    static PG_HD BigFishState &st(Ctx &c) { return game_state<BigFishState>(c); }

    // bigfish.cpp:23-29
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->timeout = 6000;
        c.h->main_width = 20;
        c.h->main_height = 20;
    }
    // bigfish.cpp:45-59
    static PG_HD void handle_agent_collision(Ctx &c, int oi) {
        Entity &obj = c.ents[oi];
        if (obj.type == FISH) {
            Entity &a = agent_of(c);
            if (obj.rx > a.rx) {
                c.h->done = 1;
            } else {
                c.h->reward += POSITIVE_REWARD;
                obj.will_erase = 1;
                a.rx += st(c).r_inc;
                a.ry += st(c).r_inc;
                st(c).fish_eaten += 1;
            }
        }
    }
    // bigfish.cpp:61-79
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        c.h->options.center_agent = 0;
        st(c).fish_eaten = 0;
        float start_r = .5;
        if (c.h->options.distribution_mode == EasyMode)
            start_r = 1;
        st(c).r_inc = (FISH_MAX_R - start_r) / FISH_QUOTA;
        Entity &a = agent_of(c);
        a.rx = start_r;
        a.ry = start_r;
        a.y = 1 + a.ry;
    }
    // bigfish.cpp:81-107
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        MT19937 &rg = *c.rng;
        if (rand_randn(rg, 10) == 1) {
            // pow is the C double overload; the float product is widened first
            float ent_r = (float)((double)(FISH_MAX_R - FISH_MIN_R) * pow((double)rand_rand01(rg), 1.4) + (double)FISH_MIN_R);
            float ent_y = rand_rand01(rg) * (h.main_height - 2 * ent_r);
            float moves_right = rand_rand01(rg) < .5;
            float ent_vx = (float)((.15 + (double)rand_rand01(rg) * .25) * (moves_right ? 1 : -1));
            float ent_x = moves_right ? -1 * ent_r : h.main_width + ent_r;
            int ei = E::add_entity(c, ent_x, ent_y, ent_vx, 0, ent_r, FISH);
            Entity &ent = c.ents[ei];
            E::choose_random_theme(c, ent);
            E::match_aspect_ratio(c, ent);
            ent.is_reflected = !moves_right;
        }
        if (st(c).fish_eaten >= FISH_QUOTA) {
            h.done = 1;
            h.reward += COMPLETION_BONUS;
            h.level_complete = 1;
        }
        if (h.action_vx > 0)
            agent_of(c).is_reflected = 0;
        if (h.action_vx < 0)
            agent_of(c).is_reflected = 1;
    }
};

}  // namespace pg
