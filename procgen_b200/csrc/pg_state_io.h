// get_state / set_state wire format (vecgame.cpp:437-457), byte-compatible with the reference:
// Game::serialize (game.cpp:170-229), BasicAbstractGame::serialize (basic-abstract-game.cpp:1169-1223),
// Entity::serialize (entity.cpp:90-131), RandGen::serialize (randgen.cpp:100-107: the libstdc++
// text form of std::mt19937), Grid::serialize (grid.h:69-73), buffer.h, and each game's tail
// (games/<name>.cpp serialize/deserialize, cited per game below).
//
// Host only. The env's records are copied out of HBM into a HostEnv, converted here, and copied
// back for set_state; nothing on the step path touches this file.
#pragma once
#include <stdexcept>
#include <string>
#include <vector>

#include "pg_kernels.cuh"

namespace pg {
namespace host {

constexpr int SERIALIZE_VERSION = 0;  // game.h
constexpr int END_OF_BUFFER = 0xCAFECAFE;  // vecgame.cpp

struct HostEnv {
    EnvHdr h;
    std::vector<Entity> ents;     // ent_cap + 1 records (the last one is the agent's ghost slot)
    std::vector<int16_t> grid;    // grid_cap cells
    MT19937 rng, lvl_rng;
    std::vector<int32_t> scratch;
    int ent_cap = 0;
};

// options that are constant per VecGame and not kept per env on the device
struct ConstGameFields {
    int use_easy_jump = 0, plain_assets = 0, physics_mode = 0, game_type = 0;
};

// ---- buffer.h
struct WriteBuf {
    char *data;
    size_t offset = 0, length;
    WriteBuf(char *d, size_t n) : data(d), length(n) {}
    void need(size_t n) {
        if (offset + n > length)
            throw std::runtime_error("state buffer too small");
    }
    void write_int(int v) {
        need(4);
        memcpy(data + offset, &v, 4);
        offset += 4;
    }
    void write_bool(bool b) { write_int(b ? 1 : 0); }
    void write_float(float f) {
        need(4);
        memcpy(data + offset, &f, 4);
        offset += 4;
    }
    void write_string(const std::string &s) {
        write_int((int)s.size());
        need(s.size());
        memcpy(data + offset, s.data(), s.size());
        offset += s.size();
    }
};
struct ReadBuf {
    const char *data;
    size_t offset = 0, length;
    ReadBuf(const char *d, size_t n) : data(d), length(n) {}
    void need(size_t n) {
        if (offset + n > length)
            throw std::runtime_error("state buffer truncated");
    }
    int read_int() {
        need(4);
        int v;
        memcpy(&v, data + offset, 4);
        offset += 4;
        return v;
    }
    bool read_bool() { return read_int() > 0; }
    float read_float() {
        need(4);
        float f;
        memcpy(&f, data + offset, 4);
        offset += 4;
        return f;
    }
    std::string read_string() {
        int n = read_int();
        if (n < 0)
            throw std::runtime_error("bad string length in state");
        need((size_t)n);
        std::string s(data + offset, (size_t)n);
        offset += (size_t)n;
        return s;
    }
};

// ---- RandGen (randgen.cpp:100-114). libstdc++ prints the 624 state words and the position,
// separated by single spaces; its state is always a fully regenerated generation, ours twists
// words on demand (pg_rng.cuh), so the words [gen, 624) are brought up to date in a copy first.
inline std::string mt_to_text(const MT19937 &src) {
    MT19937 s = src;
    if (s.p < 624) {
        for (int k = s.gen; k < 624; k++) {
            const int k1 = (k + 1 == 624) ? 0 : k + 1;
            const int km = (k + 397 >= 624) ? k + 397 - 624 : k + 397;
            const uint32_t y = (s.mt[k] & 0x80000000u) | (s.mt[k1] & 0x7fffffffu);
            s.mt[k] = s.mt[km] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
    }
    std::string out;
    out.reserve(624 * 11 + 8);
    char tmp[16];
    for (int i = 0; i < 624; i++) {
        snprintf(tmp, sizeof(tmp), "%u ", s.mt[i]);
        out += tmp;
    }
    snprintf(tmp, sizeof(tmp), "%d", s.p);
    out += tmp;
    return out;
}
inline void mt_from_text(MT19937 &s, const std::string &text) {
    const char *p = text.c_str();
    char *end = nullptr;
    for (int i = 0; i < 624; i++) {
        unsigned long v = strtoul(p, &end, 10);
        if (end == p)
            throw std::runtime_error("bad mt19937 text in state");
        s.mt[i] = (uint32_t)v;
        p = end;
    }
    long pos = strtol(p, &end, 10);
    if (end == p || pos < 0 || pos > 624)
        throw std::runtime_error("bad mt19937 position in state");
    s.p = (int32_t)pos;
    s.gen = 624;  // every word of the imported generation is already regenerated
}
inline void write_randgen(WriteBuf &b, const MT19937 &s) {
    b.write_int(s.seeded);
    b.write_string(mt_to_text(s));
}
inline void read_randgen(ReadBuf &b, MT19937 &s) {
    s.seeded = b.read_int();
    mt_from_text(s, b.read_string());
}

// ---- Entity (entity.cpp:90-165)
inline void write_entity(WriteBuf &b, const Entity &e) {
    b.write_float(e.x);
    b.write_float(e.y);
    b.write_float(e.vx);
    b.write_float(e.vy);
    b.write_float(e.rx);
    b.write_float(e.ry);
    b.write_int(e.type);
    b.write_int(e.image_type);
    b.write_int(e.image_theme);
    b.write_int(e.render_z);
    b.write_int(e.will_erase);
    b.write_int(e.collides_with_entities);
    b.write_float(e.collision_margin);
    b.write_float(e.rotation);
    b.write_float(e.vrot);
    b.write_int(e.is_reflected);
    b.write_int(e.fire_time);
    b.write_int(e.spawn_time);
    b.write_int(e.life_time);
    b.write_int(e.expire_time);
    b.write_int(e.use_abs_coords);
    b.write_float(e.friction);
    b.write_int(e.smart_step);
    b.write_int(e.avoids_collisions);
    b.write_int(e.auto_erase);
    b.write_float(e.alpha);
    b.write_float(e.health);
    b.write_float(e.theta);
    b.write_float(e.grow_rate);
    b.write_float(e.alpha_decay);
    b.write_float(e.climber_spawn_x);
}
inline void read_entity(ReadBuf &b, Entity &e) {
    memset(&e, 0, sizeof(e));
    e.x = b.read_float();
    e.y = b.read_float();
    e.vx = b.read_float();
    e.vy = b.read_float();
    e.rx = b.read_float();
    e.ry = b.read_float();
    e.type = b.read_int();
    e.image_type = b.read_int();
    e.image_theme = b.read_int();
    e.render_z = b.read_int();
    e.will_erase = (uint8_t)(b.read_int() != 0);
    e.collides_with_entities = (uint8_t)(b.read_int() != 0);
    e.collision_margin = b.read_float();
    e.rotation = b.read_float();
    e.vrot = b.read_float();
    e.is_reflected = (uint8_t)(b.read_int() != 0);
    e.fire_time = b.read_int();
    e.spawn_time = b.read_int();
    e.life_time = b.read_int();
    e.expire_time = b.read_int();
    e.use_abs_coords = (uint8_t)(b.read_int() != 0);
    e.friction = b.read_float();
    e.smart_step = (uint8_t)(b.read_int() != 0);
    e.avoids_collisions = (uint8_t)(b.read_int() != 0);
    e.auto_erase = (uint8_t)(b.read_int() != 0);
    e.alpha = b.read_float();
    e.health = b.read_float();
    e.theta = b.read_float();
    e.grow_rate = b.read_float();
    e.alpha_decay = b.read_float();
    e.climber_spawn_x = b.read_float();
}

template <class T>
T &tail(HostEnv &e) {
    return *reinterpret_cast<T *>(e.h.game_state);
}
template <class T>
const T &tail(const HostEnv &e) {
    return *reinterpret_cast<const T *>(e.h.game_state);
}

inline int find_entity_index(const HostEnv &e, int type) {  // basic-abstract-game.cpp:440-449
    int index = -1;
    for (int i = 0; i < e.h.n_ents; i++)
        if (e.ents[i].type == type)
            index = i;
    return index;
}

// ---- per-game tails
inline void write_tail(int game_id, const HostEnv &e, WriteBuf &b) {
    switch (game_id) {
    case GAME_BIGFISH: {  // bigfish.cpp:169-173
        const auto &s = tail<BigFishState>(e);
        b.write_int(s.fish_eaten);
        b.write_float(s.r_inc);
        break;
    }
    case GAME_BOSSFIGHT: {  // bossfight.cpp:415-441
        const auto &s = tail<BossfightState>(e);
        b.write_int(s.n_attack_modes);
        for (int i = 0; i < s.n_attack_modes; i++) b.write_int(s.attack_modes[i]);
        b.write_int(s.last_fire_time);
        b.write_int(s.time_to_swap);
        b.write_int(s.invulnerable_duration);
        b.write_int(s.vulnerable_duration);
        b.write_int(s.num_rounds);
        b.write_int(s.round_num);
        b.write_int(s.round_health);
        b.write_int(s.boss_vel_timeout);
        b.write_int(s.curr_vel_timeout);
        b.write_int(s.attack_mode);
        b.write_int(s.player_laser_theme);
        b.write_int(s.boss_laser_theme);
        b.write_int(s.damaged_until_time);
        b.write_bool(s.shields_are_up != 0);
        b.write_bool(s.barriers_moves_right != 0);
        b.write_float(s.base_fire_prob);
        b.write_float(s.boss_bullet_vel);
        b.write_float(s.barrier_vel);
        b.write_float(s.barrier_spawn_prob);
        b.write_float(s.rand_pct);
        b.write_float(s.rand_fire_pct);
        b.write_float(s.rand_pct_x);
        b.write_float(s.rand_pct_y);
        break;
    }
    case GAME_CAVEFLYER:  // no fields of its own
        break;
    case GAME_CHASER: {  // chaser.cpp:388-398
        const auto &s = tail<ChaserState>(e);
        const int32_t *fc = e.scratch.data() + ChaserGame::MAZE_WORDS;
        const int32_t *isv = e.scratch.data() + ChaserGame::MAZE_WORDS + ChaserGame::LIST_WORDS;
        b.write_int(s.n_free_cells);
        for (int i = 0; i < s.n_free_cells; i++) b.write_int(fc[i]);
        b.write_int(e.h.grid_size);
        for (int i = 0; i < e.h.grid_size; i++) b.write_bool(isv[i] != 0);
        b.write_int(s.eat_timeout);
        b.write_int(s.egg_timeout);
        b.write_int(s.eat_time);
        b.write_int(s.total_enemies);
        b.write_int(s.total_orbs);
        b.write_int(s.orbs_collected);
        b.write_int(s.maze_dim);
        break;
    }
    case GAME_CLIMBER: {  // climber.cpp serialize
        const auto &s = tail<ClimberState>(e);
        b.write_bool(s.has_support != 0);
        b.write_bool(s.facing_right != 0);
        b.write_int(s.coin_quota);
        b.write_int(s.coins_collected);
        b.write_int(s.wall_theme);
        b.write_float(s.gravity);
        b.write_float(s.air_control);
        break;
    }
    case GAME_COINRUN: {  // coinrun.cpp:500-509
        const auto &s = tail<CoinRunState>(e);
        b.write_float(s.last_agent_y);
        b.write_int(s.wall_theme);
        b.write_bool(s.has_support != 0);
        b.write_bool(s.facing_right != 0);
        b.write_bool(s.is_on_crate != 0);
        b.write_float(s.gravity);
        b.write_float(s.air_control);
        break;
    }
    case GAME_DODGEBALL: {  // dodgeball.cpp:442-451
        const auto &s = tail<DodgeballState>(e);
        b.write_float(s.min_dim);
        b.write_float(s.hard_min_dim);
        b.write_float(s.ball_vscale);
        b.write_float(s.ball_r);
        b.write_int(s.last_fire_time);
        b.write_int(s.num_enemies);
        b.write_int(s.enemy_fire_delay);
        break;
    }
    case GAME_FRUITBOT: {  // fruitbot.cpp serialize
        const auto &s = tail<FruitBotState>(e);
        b.write_float(s.min_dim);
        b.write_float(s.bullet_vscale);
        b.write_int(s.last_fire_time);
        break;
    }
    case GAME_HEIST: {  // heist.cpp:211-216
        const auto &s = tail<HeistState>(e);
        b.write_int(s.num_keys);
        b.write_int(s.world_dim);
        b.write_int(s.num_keys);
        for (int i = 0; i < s.num_keys; i++) b.write_bool(s.has_keys[i] != 0);
        break;
    }
    case GAME_JUMPER: {  // jumper.cpp:445-454
        const auto &s = tail<JumperState>(e);
        b.write_int(s.jump_count);
        b.write_int(s.jump_delta);
        b.write_int(s.jump_time);
        b.write_bool(s.has_support != 0);
        b.write_bool(s.facing_right != 0);
        b.write_int(s.wall_theme);
        b.write_float(s.compass_dim);
        break;
    }
    case GAME_LEAPER: {  // leaper.cpp serialize
        const auto &s = tail<LeaperState>(e);
        b.write_int(s.bottom_road_y);
        b.write_int(s.n_road);
        for (int i = 0; i < s.n_road; i++) b.write_float(s.road_lane_speeds[i]);
        b.write_int(s.bottom_water_y);
        b.write_int(s.n_water);
        for (int i = 0; i < s.n_water; i++) b.write_float(s.water_lane_speeds[i]);
        b.write_int(s.goal_y);
        break;
    }
    case GAME_MAZE: {  // maze.cpp serialize
        const auto &s = tail<MazeState>(e);
        b.write_int(s.maze_dim);
        b.write_int(s.world_dim);
        break;
    }
    case GAME_MINER: {  // miner.cpp serialize
        b.write_int(tail<MinerState>(e).diamonds_remaining);
        break;
    }
    case GAME_NINJA: {  // ninja.cpp serialize
        const auto &s = tail<NinjaState>(e);
        b.write_bool(s.has_support != 0);
        b.write_bool(s.facing_right != 0);
        b.write_int(s.last_fire_time);
        b.write_int(s.wall_theme);
        b.write_float(s.gravity);
        b.write_float(s.air_control);
        b.write_float(s.jump_charge);
        b.write_float(s.jump_charge_inc);
        break;
    }
    case GAME_PLUNDER: {  // plunder.cpp:243-259
        const auto &s = tail<PlunderState>(e);
        b.write_int(s.last_fire_time);
        b.write_int(s.num_lanes);
        for (int i = 0; i < s.num_lanes; i++) b.write_bool(s.lane_directions[i] != 0);
        b.write_int(6);
        for (int i = 0; i < 6; i++) b.write_bool(s.target_bools[i] != 0);
        b.write_int(6);
        for (int i = 0; i < 6; i++) b.write_int(s.image_permutation[i]);
        b.write_int(s.num_lanes);
        for (int i = 0; i < s.num_lanes; i++) b.write_float(s.lane_vels[i]);
        b.write_int(s.num_lanes);
        b.write_int(s.num_current_ship_types);
        b.write_int(s.targets_hit);
        b.write_int(s.target_quota);
        b.write_float(s.juice_left);
        b.write_float(s.r_scale);
        b.write_float(s.spawn_prob);
        b.write_float(s.legend_r);
        b.write_float(s.min_agent_x);
        break;
    }
    case GAME_STARPILOT: {  // starpilot.cpp:451-454: the remaining spawners, in list order
        const auto &s = tail<StarpilotState>(e);
        const Entity *recs = reinterpret_cast<const Entity *>(e.scratch.data());
        const int32_t *order = e.scratch.data() + StarpilotGame::MAX_SPAWNERS * StarpilotGame::ENT_WORDS;
        b.write_int(s.n_spawners);
        for (int i = 0; i < s.n_spawners; i++) write_entity(b, recs[order[i]]);
        break;
    }
    default:
        throw std::runtime_error("get_state: unknown game id");
    }
}

// The device keeps a few derived values the reference recomputes or keeps outside the blob
// (entity indices instead of shared_ptrs, list lengths); deserialize restores those too.
inline void read_tail(int game_id, HostEnv &e, ReadBuf &b) {
    switch (game_id) {
    case GAME_BIGFISH: {
        auto &s = tail<BigFishState>(e);
        s.fish_eaten = b.read_int();
        s.r_inc = b.read_float();
        break;
    }
    case GAME_BOSSFIGHT: {  // bossfight.cpp:443-476
        auto &s = tail<BossfightState>(e);
        s.n_attack_modes = b.read_int();
        if (s.n_attack_modes < 0 || s.n_attack_modes > 8)
            throw std::runtime_error("set_state: bad attack_modes length");
        for (int i = 0; i < s.n_attack_modes; i++) s.attack_modes[i] = b.read_int();
        s.last_fire_time = b.read_int();
        s.time_to_swap = b.read_int();
        s.invulnerable_duration = b.read_int();
        s.vulnerable_duration = b.read_int();
        s.num_rounds = b.read_int();
        s.round_num = b.read_int();
        s.round_health = b.read_int();
        s.boss_vel_timeout = b.read_int();
        s.curr_vel_timeout = b.read_int();
        s.attack_mode = b.read_int();
        s.player_laser_theme = b.read_int();
        s.boss_laser_theme = b.read_int();
        s.damaged_until_time = b.read_int();
        s.shields_are_up = b.read_bool();
        s.barriers_moves_right = b.read_bool();
        s.base_fire_prob = b.read_float();
        s.boss_bullet_vel = b.read_float();
        s.barrier_vel = b.read_float();
        s.barrier_spawn_prob = b.read_float();
        s.rand_pct = b.read_float();
        s.rand_fire_pct = b.read_float();
        s.rand_pct_x = b.read_float();
        s.rand_pct_y = b.read_float();
        s.boss_idx = find_entity_index(e, BossfightGame::BOSS);
        s.shields_idx = find_entity_index(e, BossfightGame::SHIELDS);
        if (s.boss_idx < 0 || s.shields_idx < 0)
            throw std::runtime_error("set_state: bossfight state without boss or shields");
        break;
    }
    case GAME_CAVEFLYER:
        break;
    case GAME_CHASER: {
        auto &s = tail<ChaserState>(e);
        int32_t *fc = e.scratch.data() + ChaserGame::MAZE_WORDS;
        int32_t *isv = e.scratch.data() + ChaserGame::MAZE_WORDS + ChaserGame::LIST_WORDS;
        s.n_free_cells = b.read_int();
        if (s.n_free_cells < 0 || s.n_free_cells > ChaserGame::LIST_WORDS)
            throw std::runtime_error("set_state: bad free_cells length");
        for (int i = 0; i < s.n_free_cells; i++) fc[i] = b.read_int();
        int n = b.read_int();
        if (n < 0 || n > ChaserGame::LIST_WORDS)
            throw std::runtime_error("set_state: bad is_space_vec length");
        for (int i = 0; i < n; i++) isv[i] = b.read_bool();
        s.eat_timeout = b.read_int();
        s.egg_timeout = b.read_int();
        s.eat_time = b.read_int();
        s.total_enemies = b.read_int();
        s.total_orbs = b.read_int();
        s.orbs_collected = b.read_int();
        s.maze_dim = b.read_int();
        break;
    }
    case GAME_CLIMBER: {
        auto &s = tail<ClimberState>(e);
        s.has_support = b.read_bool();
        s.facing_right = b.read_bool();
        s.coin_quota = b.read_int();
        s.coins_collected = b.read_int();
        s.wall_theme = b.read_int();
        s.gravity = b.read_float();
        s.air_control = b.read_float();
        break;
    }
    case GAME_COINRUN: {
        auto &s = tail<CoinRunState>(e);
        s.last_agent_y = b.read_float();
        s.wall_theme = b.read_int();
        s.has_support = b.read_bool();
        s.facing_right = b.read_bool();
        s.is_on_crate = b.read_bool();
        s.gravity = b.read_float();
        s.air_control = b.read_float();
        break;
    }
    case GAME_DODGEBALL: {
        auto &s = tail<DodgeballState>(e);
        s.min_dim = b.read_float();
        s.hard_min_dim = b.read_float();
        s.ball_vscale = b.read_float();
        s.ball_r = b.read_float();
        s.last_fire_time = b.read_int();
        s.num_enemies = b.read_int();
        s.enemy_fire_delay = b.read_int();
        break;
    }
    case GAME_FRUITBOT: {
        auto &s = tail<FruitBotState>(e);
        s.min_dim = b.read_float();
        s.bullet_vscale = b.read_float();
        s.last_fire_time = b.read_int();
        break;
    }
    case GAME_HEIST: {
        auto &s = tail<HeistState>(e);
        s.num_keys = b.read_int();
        s.world_dim = b.read_int();
        int n = b.read_int();
        if (n < 0 || n > 4)
            throw std::runtime_error("set_state: bad has_keys length");
        for (int i = 0; i < n; i++) s.has_keys[i] = b.read_bool();
        break;
    }
    case GAME_JUMPER: {  // jumper.cpp:456-469
        auto &s = tail<JumperState>(e);
        s.jump_count = b.read_int();
        s.jump_delta = b.read_int();
        s.jump_time = b.read_int();
        s.has_support = b.read_bool();
        s.facing_right = b.read_bool();
        s.wall_theme = b.read_int();
        s.compass_dim = b.read_float();
        s.goal_idx = find_entity_index(e, JumperGame::GOAL);
        if (s.goal_idx < 0)
            throw std::runtime_error("set_state: jumper state without a goal");
        break;
    }
    case GAME_LEAPER: {
        auto &s = tail<LeaperState>(e);
        s.bottom_road_y = b.read_int();
        s.n_road = b.read_int();
        if (s.n_road < 0 || s.n_road > 8)
            throw std::runtime_error("set_state: bad road_lane_speeds length");
        for (int i = 0; i < s.n_road; i++) s.road_lane_speeds[i] = b.read_float();
        s.bottom_water_y = b.read_int();
        s.n_water = b.read_int();
        if (s.n_water < 0 || s.n_water > 8)
            throw std::runtime_error("set_state: bad water_lane_speeds length");
        for (int i = 0; i < s.n_water; i++) s.water_lane_speeds[i] = b.read_float();
        s.goal_y = b.read_int();
        break;
    }
    case GAME_MAZE: {
        auto &s = tail<MazeState>(e);
        s.maze_dim = b.read_int();
        s.world_dim = b.read_int();
        break;
    }
    case GAME_MINER:
        tail<MinerState>(e).diamonds_remaining = b.read_int();
        break;
    case GAME_NINJA: {
        auto &s = tail<NinjaState>(e);
        s.has_support = b.read_bool();
        s.facing_right = b.read_bool();
        s.last_fire_time = b.read_int();
        s.wall_theme = b.read_int();
        s.gravity = b.read_float();
        s.air_control = b.read_float();
        s.jump_charge = b.read_float();
        s.jump_charge_inc = b.read_float();
        break;
    }
    case GAME_PLUNDER: {
        auto &s = tail<PlunderState>(e);
        s.last_fire_time = b.read_int();
        int n = b.read_int();
        if (n < 0 || n > 5)
            throw std::runtime_error("set_state: bad lane_directions length");
        for (int i = 0; i < n; i++) s.lane_directions[i] = b.read_bool();
        n = b.read_int();
        if (n < 0 || n > 6)
            throw std::runtime_error("set_state: bad target_bools length");
        for (int i = 0; i < n; i++) s.target_bools[i] = b.read_bool();
        n = b.read_int();
        if (n < 0 || n > 6)
            throw std::runtime_error("set_state: bad image_permutation length");
        for (int i = 0; i < n; i++) s.image_permutation[i] = b.read_int();
        n = b.read_int();
        if (n < 0 || n > 5)
            throw std::runtime_error("set_state: bad lane_vels length");
        for (int i = 0; i < n; i++) s.lane_vels[i] = b.read_float();
        s.num_lanes = b.read_int();
        s.num_current_ship_types = b.read_int();
        s.targets_hit = b.read_int();
        s.target_quota = b.read_int();
        s.juice_left = b.read_float();
        s.r_scale = b.read_float();
        s.spawn_prob = b.read_float();
        s.legend_r = b.read_float();
        s.min_agent_x = b.read_float();
        break;
    }
    case GAME_STARPILOT: {  // starpilot.cpp:456-461 (init_hps is replayed by the caller on the device side: it
                            // depends only on the distribution mode, whose values are already in the tail)
        auto &s = tail<StarpilotState>(e);
        Entity *recs = reinterpret_cast<Entity *>(e.scratch.data());
        int32_t *order = e.scratch.data() + StarpilotGame::MAX_SPAWNERS * StarpilotGame::ENT_WORDS;
        int n = b.read_int();
        if (n < 0 || n > StarpilotGame::MAX_SPAWNERS)
            throw std::runtime_error("set_state: bad spawner count");
        for (int i = 0; i < n; i++) {
            read_entity(b, recs[i]);
            order[i] = i;
        }
        s.n_spawners = n;
        break;
    }
    default:
        throw std::runtime_error("set_state: unknown game id");
    }
}

// ---- Game + BasicAbstractGame
inline void serialize_env(const char *game_name, int game_id, const HostEnv &e, const ConstGameFields &cf, WriteBuf &b) {
    const EnvHdr &h = e.h;
    b.write_int(SERIALIZE_VERSION);
    b.write_string(game_name);
    b.write_int(h.options.paint_vel_info);
    b.write_int(h.options.use_generated_assets);
    b.write_int(h.options.use_monochrome_assets);
    b.write_int(h.options.restrict_themes);
    b.write_int(h.options.use_backgrounds);
    b.write_int(h.options.center_agent);
    b.write_int(h.options.debug_mode);
    b.write_int(h.options.distribution_mode);
    b.write_int(h.options.use_sequential_levels);
    b.write_int(cf.use_easy_jump);
    b.write_int(cf.plain_assets);
    b.write_int(cf.physics_mode);
    b.write_int(h.grid_step);
    b.write_int(h.level_seed_low);
    b.write_int(h.level_seed_high);
    b.write_int(cf.game_type);
    b.write_int(h.game_n);
    write_randgen(b, e.lvl_rng);
    write_randgen(b, e.rng);
    b.write_float(h.reward);
    b.write_int(h.done);
    b.write_int(h.level_complete);
    b.write_int(h.action);
    b.write_int(h.timeout);
    b.write_int(h.current_level_seed);
    b.write_int(h.prev_level_seed);
    b.write_int(h.episodes_remaining);
    b.write_int(h.episode_done);
    b.write_int(h.last_reward_timer);
    b.write_float(h.last_reward);
    b.write_int(h.default_action);
    b.write_int(h.fixed_asset_seed);
    b.write_int(h.cur_time);
    b.write_int(0);  // is_waiting_for_step: get_state waits for the step first
    // BasicAbstractGame
    b.write_int(h.grid_size);
    if (h.agent_idx >= e.ent_cap)
        throw std::runtime_error("get_state: the agent is not in the entity list");
    b.write_int(h.n_ents);
    for (int i = 0; i < h.n_ents; i++) write_entity(b, e.ents[i]);
    b.write_int(0);  // use_procgen_background: every game loads real backgrounds (basic-abstract-game.cpp:54-66)
    b.write_int(h.background_index);
    b.write_float(h.bg_tile_ratio);
    b.write_float(h.bg_pct_x);
    b.write_float(h.char_dim);
    b.write_int(h.last_move_action);
    b.write_int(h.move_action);
    b.write_int(h.special_action);
    b.write_float(h.mixrate);
    b.write_float(h.maxspeed);
    b.write_float(h.max_jump);
    b.write_float(h.action_vx);
    b.write_float(h.action_vy);
    b.write_float(h.action_vrot);
    b.write_float(h.center_x);
    b.write_float(h.center_y);
    b.write_int(h.random_agent_start);
    b.write_int(h.has_useful_vel_info);
    b.write_int(h.step_rand_int);
    {
        // asset_rand_gen: only the generated-asset path ever seeds or draws from it, so it is the
        // default-constructed engine (std::mt19937 default seed 5489) for the whole run
        MT19937 asset;
        memset(&asset, 0, sizeof(asset));
        mt_seed(asset, 5489u);
        asset.seeded = 0;
        write_randgen(b, asset);
    }
    b.write_int(h.main_width);
    b.write_int(h.main_height);
    b.write_int(h.out_of_bounds_object);
    b.write_float(h.unit);
    b.write_float(h.view_dim);
    b.write_float(h.x_off);
    b.write_float(h.y_off);
    b.write_float(h.visibility);
    b.write_float(h.min_visibility);
    // Grid<int>::serialize (grid.h:69-73)
    b.write_int(h.main_width);
    b.write_int(h.main_height);
    const int cells = h.main_width * h.main_height;
    b.write_int(cells);
    for (int i = 0; i < cells; i++) b.write_int((int)e.grid[i]);
    write_tail(game_id, e, b);
    b.write_int(END_OF_BUFFER);
}

inline void deserialize_env(const char *game_name, int game_id, HostEnv &e, ReadBuf &b) {
    EnvHdr &h = e.h;
    if (b.read_int() != SERIALIZE_VERSION)
        throw std::runtime_error("set_state: serialize version mismatch");
    if (b.read_string() != game_name)
        throw std::runtime_error("set_state: state belongs to another game");
    h.options.paint_vel_info = (uint8_t)b.read_int();
    h.options.use_generated_assets = (uint8_t)b.read_int();
    h.options.use_monochrome_assets = (uint8_t)b.read_int();
    h.options.restrict_themes = (uint8_t)b.read_int();
    h.options.use_backgrounds = (uint8_t)b.read_int();
    h.options.center_agent = (uint8_t)b.read_int();
    h.options.debug_mode = b.read_int();
    h.options.distribution_mode = b.read_int();
    h.options.use_sequential_levels = (uint8_t)b.read_int();
    if (h.options.use_generated_assets)
        throw std::runtime_error("set_state: use_generated_assets is not supported");
    b.read_int();  // use_easy_jump   (per-VecGame constants: kept as constructed)
    b.read_int();  // plain_assets
    b.read_int();  // physics_mode
    h.grid_step = b.read_int();
    h.level_seed_low = b.read_int();
    h.level_seed_high = b.read_int();
    b.read_int();  // game_type
    h.game_n = b.read_int();
    read_randgen(b, e.lvl_rng);
    read_randgen(b, e.rng);
    h.reward = b.read_float();
    h.done = b.read_int();
    h.level_complete = b.read_int();
    h.action = b.read_int();
    h.timeout = b.read_int();
    h.current_level_seed = b.read_int();
    h.prev_level_seed = b.read_int();
    h.episodes_remaining = b.read_int();
    h.episode_done = b.read_int();
    h.last_reward_timer = b.read_int();
    h.last_reward = b.read_float();
    h.default_action = b.read_int();
    h.fixed_asset_seed = b.read_int();
    h.cur_time = b.read_int();
    b.read_int();  // is_waiting_for_step
    h.grid_size = b.read_int();
    const int n_ents = b.read_int();
    if (n_ents < 0 || n_ents > e.ent_cap)
        throw std::runtime_error("set_state: more entities than this build's capacity for the game");
    for (int i = 0; i < n_ents; i++) read_entity(b, e.ents[i]);
    h.n_ents = n_ents;
    if (h.max_ents_seen < n_ents)
        h.max_ents_seen = n_ents;
    h.agent_idx = find_entity_index(e, PLAYER);  // basic-abstract-game.cpp:1231-1233
    if (h.agent_idx < 0)
        throw std::runtime_error("set_state: state without an agent");
    b.read_int();  // use_procgen_background
    h.background_index = b.read_int();
    h.bg_tile_ratio = b.read_float();
    h.bg_pct_x = b.read_float();
    h.char_dim = b.read_float();
    h.last_move_action = b.read_int();
    h.move_action = b.read_int();
    h.special_action = b.read_int();
    h.mixrate = b.read_float();
    h.maxspeed = b.read_float();
    h.max_jump = b.read_float();
    h.action_vx = b.read_float();
    h.action_vy = b.read_float();
    h.action_vrot = b.read_float();
    h.center_x = b.read_float();
    h.center_y = b.read_float();
    h.random_agent_start = b.read_int();
    h.has_useful_vel_info = b.read_int();
    h.step_rand_int = b.read_int();
    {
        MT19937 asset;
        read_randgen(b, asset);  // asset_rand_gen: unused without generated assets
    }
    h.main_width = b.read_int();
    h.main_height = b.read_int();
    h.out_of_bounds_object = b.read_int();
    h.unit = b.read_float();
    h.view_dim = b.read_float();
    h.x_off = b.read_float();
    h.y_off = b.read_float();
    h.visibility = b.read_float();
    h.min_visibility = b.read_float();
    const int gw = b.read_int(), gh = b.read_int();
    const int cells = b.read_int();
    if (gw != h.main_width || gh != h.main_height || cells != gw * gh || cells > (int)e.grid.size())
        throw std::runtime_error("set_state: grid does not fit this build's capacity for the game");
    for (int i = 0; i < cells; i++) e.grid[i] = (int16_t)b.read_int();
    read_tail(game_id, e, b);
    if (b.read_int() != END_OF_BUFFER)
        throw std::runtime_error("set_state: trailing bytes in state");
}

}  // namespace host
}  // namespace pg
