// Per-environment state as it lives in HBM.
//
// One env = four fixed-size records in four arrays indexed by env (structure-of-arrays across
// record kinds, contiguous within a record so a CTA can move its env with bulk copies):
//   EnvHdr     scalars of Game + BasicAbstractGame + the per-game tail   (game.h:64-106,
//              basic-abstract-game.h:115-159)
//   Entity[]   ordered entity list, 128 B per entity (entity.h:9-48)
//   int16[]    grid, row-major y*w+x, y up (grid.h:40-62); ids >= 1000 exist (chaser ORB) so int16
//   MT19937 x2 rand_gen and level_seed_rand_gen (game.h:77-78)
#pragma once
#include "pg_common.cuh"
#include "pg_rng.cuh"

namespace pg {

// entity.h:9-48. Field order groups what the physics loop touches into the first 64 B.
struct alignas(16) Entity {
    float x, y, vx, vy, rx, ry;
    int32_t type;
    int32_t image_type;
    int32_t image_theme;
    int32_t render_z;
    float collision_margin;
    float rotation;
    float vrot;
    uint8_t will_erase;
    uint8_t collides_with_entities;
    uint8_t is_reflected;
    uint8_t use_abs_coords;
    uint8_t smart_step;
    uint8_t avoids_collisions;
    uint8_t auto_erase;
    uint8_t pad0;
    // ---- 64 B
    int32_t fire_time;
    int32_t spawn_time;
    int32_t life_time;
    int32_t expire_time;
    float friction;
    float alpha;
    float health;
    float theta;
    float grow_rate;
    float alpha_decay;
    float climber_spawn_x;
    int32_t pad1[5];
};
static_assert(sizeof(Entity) == 128, "Entity must be one 128-byte line");

// GameOptions (game.h:45-62) — per env because games overwrite center_agent in game_reset.
struct Options {
    uint8_t paint_vel_info;
    uint8_t use_generated_assets;
    uint8_t use_monochrome_assets;
    uint8_t restrict_themes;
    uint8_t use_backgrounds;
    uint8_t center_agent;
    uint8_t use_sequential_levels;
    uint8_t pad;
    int32_t debug_mode;
    int32_t distribution_mode;
};

constexpr int GAME_STATE_BYTES = 256;

struct alignas(16) EnvHdr {
    // ---- Game (game.h:64-106)
    Options options;
    int32_t game_id;
    int32_t game_n;
    int32_t grid_step;
    int32_t level_seed_low;
    int32_t level_seed_high;
    float reward;          // step_data.reward
    int32_t done;          // step_data.done
    int32_t level_complete;
    int32_t action;
    int32_t timeout;
    int32_t current_level_seed;
    int32_t prev_level_seed;
    int32_t episodes_remaining;
    int32_t episode_done;
    int32_t last_reward_timer;
    float last_reward;
    int32_t default_action;
    int32_t fixed_asset_seed;
    int32_t cur_time;
    int32_t reset_count;
    float total_reward;
    int32_t initial_reset_complete;
    // ---- BasicAbstractGame (basic-abstract-game.h:115-159)
    int32_t grid_size;
    int32_t n_ents;
    int32_t agent_idx;      // index of the agent in the entity list; == ent_cap when the agent was
                            // erased this step and lives on in the ghost slot (shared_ptr semantics)
    int32_t background_index;
    float bg_tile_ratio;
    float bg_pct_x;
    float char_dim;
    int32_t last_move_action;
    int32_t move_action;
    int32_t special_action;
    float mixrate;
    float maxspeed;
    float max_jump;
    float action_vx;
    float action_vy;
    float action_vrot;
    float center_x;
    float center_y;
    int32_t random_agent_start;
    int32_t has_useful_vel_info;
    int32_t step_rand_int;
    int32_t main_width;
    int32_t main_height;
    int32_t out_of_bounds_object;
    float unit;
    float view_dim;
    float x_off;
    float y_off;
    float visibility;
    float min_visibility;
    uint32_t err;           // ErrBits, sticky
    int32_t max_ents_seen;
    int32_t max_blits_seen; // high-water marks of the frame builder (capacity planning; not game state)
    int32_t max_rots_seen;
#ifdef PG_PHASE_TIMING
    uint32_t dbg_phase[12];  // profiling variant only: SM cycles spent in marked phases of the last reset
#endif
    // ---- per-game tail (the fields each games/*.cpp class adds)
    alignas(8) unsigned char game_state[GAME_STATE_BYTES];
};

// Asset metadata that game LOGIC needs (basic-abstract-game.cpp:79-123, 1014-1046) plus where
// each sprite lives in the device atlas. Built on the host at init; constant afterwards.
struct SpriteDesc {
    uint32_t off;   // offset in 32-bit texels into the atlas
    uint16_t w, h;  // 0,0 = no asset for this (type, theme)
};

constexpr int MAX_BACKGROUNDS = 64;

struct GameAssets {
    SpriteDesc sprites[MAX_ASSETS * MAX_IMAGE_THEMES];   // [type + 100*theme], ARGB32 premultiplied
    float aspect[MAX_ASSETS * MAX_IMAGE_THEMES];         // float(width * 1.0 / height)
    int32_t num_themes[MAX_ASSETS];                      // asset_num_themes
    SpriteDesc backgrounds[MAX_BACKGROUNDS];             // RGB32
    int32_t num_backgrounds;
    int32_t pad[3];
    int16_t sprite_slot[MAX_ASSETS * MAX_IMAGE_THEMES];  // row of the pre-scaled tile table (pg_raster.cuh TileTable), -1 = none
};

// Handle a thread uses to reach one env. Pointers are generic (global or shared).
struct Blit;
struct Ctx {
    EnvHdr *h;
    Entity *ents;
    int16_t *grid;
    MT19937 *rng;
    MT19937 *lvl_rng;
    const GameAssets *assets;
    int32_t *scratch;     // per-env level-generation workspace
    int32_t ent_cap;      // list capacity; slot [ent_cap] is the agent ghost slot
    int32_t grid_cap;
    int32_t scratch_cap;  // in int32 words
    void *rot_scratch_raw; // per-env slice for rotated-sprite / span records (setup + render kernels)
    struct Blit *blit_list; // per-env blit list the setup kernel fills and the render kernel paints
    struct Blit *cell_spill; // per-env general cell blits (setup kernel writes, render kernel reads)
    // register-resident copies of header scalars the physics loop reads constantly; refreshed by
    // ctx_refresh() whenever a game changes them (world size is chosen per episode)
    int32_t mw, mh, oob;
    // while step_entities runs: entities at or beyond this index can never block or reflect
    // anything (Defaults::may_be_obstacle); -1 = not known, scan the whole list
    int32_t obst_hi;
};

// Profiling variant (-DPG_PHASE_TIMING): PG_PHASE_BEGIN(c) ... PG_PHASE_END(c, id) accumulate cycles.
#if defined(PG_PHASE_TIMING) && defined(__CUDA_ARCH__)
#define PG_PHASE_BEGIN(c) long long _pg_t0 = clock64()
#define PG_PHASE_END(c, id)                                              \
    do {                                                                 \
        long long _pg_t1 = clock64();                                    \
        (c).h->dbg_phase[id] += (uint32_t)(_pg_t1 - _pg_t0);             \
        _pg_t0 = _pg_t1;                                                 \
    } while (0)
#define PG_PHASE_RESET(c)                                                \
    do {                                                                 \
        for (int _i = 0; _i < 12; _i++) (c).h->dbg_phase[_i] = 0;        \
    } while (0)
#define PG_PHASE_NOTE(c, id, v) (c).h->dbg_phase[id] = (uint32_t)(v)
#else
#define PG_PHASE_NOTE(c, id, v) do { } while (0)
#define PG_PHASE_BEGIN(c) do { } while (0)
#define PG_PHASE_END(c, id) do { } while (0)
#define PG_PHASE_RESET(c) do { } while (0)
#endif

PG_HD void ctx_refresh(Ctx &c) {
    c.mw = c.h->main_width;
    c.mh = c.h->main_height;
    c.oob = c.h->out_of_bounds_object;
}

template <class T>
PG_HD T &game_state(Ctx &c) {
    static_assert(sizeof(T) <= GAME_STATE_BYTES, "per-game state too large");
    return *reinterpret_cast<T *>(c.h->game_state);
}

}  // namespace pg
