// Maze on the device engine. Behaviour restated from games/maze.cpp (cited per function).
#pragma once
#include "../pg_mazegen.cuh"
#include "../pg_raster.cuh"

namespace pg {

struct MazeState {
    int32_t maze_dim;
    int32_t world_dim;
};

struct MazeGame : Defaults<MazeGame>, DrawDefaults<MazeGame> {
    using E = Engine<MazeGame>;
    static constexpr int ENT_CAP = 8;
    static constexpr int GRID_CAP = 31 * 31;
    static constexpr int SCRATCH_WORDS = 12288;  // MazeGen::words_needed(31) = 11761
    static constexpr int MAX_VISIBLE_ENTS = 64;
    static constexpr int MAX_ROT_BLITS = 0;
    static constexpr int MAX_VIEW_CELLS = 25;    // hard: whole 25x25 world; memory mode is centred (11)
    static constexpr const char *NAME = "maze";
    // is_blocked / is_blocked_ents / will_reflect are the engine defaults here: only an entity typed WALL_OBJ or as the out-of-bounds object could block
    static PG_HD bool may_be_obstacle(Ctx &c, int t) { return t == WALL_OBJ || t == c.oob; }
    static PG_HD bool may_block_or_reflect(Ctx &c, int src, int t) { return may_be_obstacle(c, t); }

    static constexpr float REWARD = 10.0;
    static constexpr int GOAL = 2;

    static PG_HD MazeState &st(Ctx &c) { return game_state<MazeState>(c); }

    // maze.cpp:16-24
    static PG_HD void init_constants(Ctx &c) {
        base_init_constants(c);
        c.h->timeout = 500;
        c.h->random_agent_start = 0;
        c.h->has_useful_vel_info = 0;
        c.h->out_of_bounds_object = WALL_OBJ;
        c.h->visibility = 8.0;
    }
    // maze.cpp:40-55
    static PG_HD void choose_world_dim(Ctx &c) {
        int dist_diff = c.h->options.distribution_mode;
        if (dist_diff == EasyMode)
            st(c).world_dim = 15;
        else if (dist_diff == HardMode)
            st(c).world_dim = 25;
        else if (dist_diff == MemoryMode)
            st(c).world_dim = 31;
        c.h->main_width = st(c).world_dim;
        c.h->main_height = st(c).world_dim;
    }
    // maze.cpp:57-99
    static PG_HD void game_reset(Ctx &c) {
        E::basic_game_reset(c);
        EnvHdr &h = *c.h;
        h.grid_step = 1;
        const int world_dim = st(c).world_dim;
        const int maze_dim = rand_randn(*c.rng, (world_dim - 1) / 2) * 2 + 3;
        st(c).maze_dim = maze_dim;
        int margin = (world_dim - maze_dim) / 2;
        MazeGen mg;
        mg.init(c, maze_dim);
        h.options.center_agent = h.options.distribution_mode == MemoryMode;
        Entity &a = agent_of(c);
        a.rx = .5;
        a.ry = .5;
        a.x = (float)(margin + .5);
        a.y = (float)(margin + .5);
        mg.generate_maze();
        mg.place_objects(GOAL, 1);
        {
            int16_t *g = c.grid;
            pg_warp_for(h.grid_size, [=](int i) { g[i] = (int16_t)WALL_OBJ; });
        }
        for (int i = 0; i < mg.maze_dim; i++)
            for (int j = 0; j < mg.maze_dim; j++)
                E::set_obj(c, margin + i, margin + j, mg.grid_get(i + MAZE_OFFSET, j + MAZE_OFFSET));
        if (margin > 0) {
            for (int i = 0; i < maze_dim + 2; i++) {
                E::set_obj(c, margin - 1, margin + i - 1, WALL_OBJ);
                E::set_obj(c, margin + maze_dim, margin + i - 1, WALL_OBJ);
                E::set_obj(c, margin + i - 1, margin - 1, WALL_OBJ);
                E::set_obj(c, margin + i - 1, margin + maze_dim, WALL_OBJ);
            }
        }
    }
    // maze.cpp:101-105
    static PG_HD void set_action_xy(Ctx &c, int move_action) {
        Defaults<MazeGame>::set_action_xy(c, move_action);
        if (c.h->action_vx != 0)
            c.h->action_vy = 0;
    }
    // maze.cpp:107-125
    static PG_HD void game_step(Ctx &c) {
        E::basic_game_step(c);
        EnvHdr &h = *c.h;
        Entity &a = agent_of(c);
        if (h.action_vx > 0)
            a.is_reflected = 1;
        if (h.action_vx < 0)
            a.is_reflected = 0;
        int ix = int(a.x);
        int iy = int(a.y);
        if (E::get_obj(c, ix, iy) == GOAL) {
            E::set_obj(c, ix, iy, SPACE);
            h.reward += REWARD;
            h.level_complete = 1;
        }
        h.done = h.reward > 0;
    }
};

}  // namespace pg
