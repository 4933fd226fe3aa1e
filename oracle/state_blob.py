"""TEST INFRASTRUCTURE — parser for the reference's get_state wire format (vecgame.cpp:437-445;
Game::serialize game.cpp:170-229; BasicAbstractGame::serialize basic-abstract-game.cpp:1169-1223;
Entity::serialize entity.cpp:90-131; RandGen::serialize randgen.cpp:100-107). Used to compare full
engine state (entities, scalars, grid) against the device, not only rewards and pixels."""
import struct

ENTITY_FIELDS = [("x", "f"), ("y", "f"), ("vx", "f"), ("vy", "f"), ("rx", "f"), ("ry", "f"), ("type", "i"), ("image_type", "i"),
                 ("image_theme", "i"), ("render_z", "i"), ("will_erase", "i"), ("collides_with_entities", "i"),
                 ("collision_margin", "f"), ("rotation", "f"), ("vrot", "f"), ("is_reflected", "i"), ("fire_time", "i"),
                 ("spawn_time", "i"), ("life_time", "i"), ("expire_time", "i"), ("use_abs_coords", "i"), ("friction", "f"),
                 ("smart_step", "i"), ("avoids_collisions", "i"), ("auto_erase", "i"), ("alpha", "f"), ("health", "f"),
                 ("theta", "f"), ("grow_rate", "f"), ("alpha_decay", "f"), ("climber_spawn_x", "f")]


class Reader:
    def __init__(self, data):
        self.d = data
        self.o = 0

    def i(self):
        v = struct.unpack_from("<i", self.d, self.o)[0]
        self.o += 4
        return v

    def f(self):
        v = struct.unpack_from("<f", self.d, self.o)[0]
        self.o += 4
        return v

    def s(self):
        n = self.i()
        v = self.d[self.o:self.o + n]
        self.o += n
        return v


def parse(blob):
    r = Reader(blob)
    out = {"version": r.i(), "game_name": r.s().decode()}
    for k in ["paint_vel_info", "use_generated_assets", "use_monochrome_assets", "restrict_themes", "use_backgrounds",
              "center_agent", "debug_mode", "distribution_mode", "use_sequential_levels", "use_easy_jump", "plain_assets",
              "physics_mode", "grid_step", "level_seed_low", "level_seed_high", "game_type", "game_n"]:
        out[k] = r.i()
    for name in ["level_seed_rand_gen", "rand_gen"]:
        seeded = r.i()
        out[name] = (seeded, r.s())
    out["reward"] = r.f()
    for k in ["done", "level_complete", "action", "timeout", "current_level_seed", "prev_level_seed", "episodes_remaining",
              "episode_done", "last_reward_timer"]:
        out[k] = r.i()
    out["last_reward"] = r.f()
    for k in ["default_action", "fixed_asset_seed", "cur_time", "is_waiting_for_step", "grid_size"]:
        out[k] = r.i()
    n = r.i()
    ents = []
    for _ in range(n):
        ents.append({name: (r.f() if t == "f" else r.i()) for name, t in ENTITY_FIELDS})
    out["entities"] = ents
    out["use_procgen_background"] = r.i()
    out["background_index"] = r.i()
    out["bg_tile_ratio"] = r.f()
    out["bg_pct_x"] = r.f()
    out["char_dim"] = r.f()
    for k in ["last_move_action", "move_action", "special_action"]:
        out[k] = r.i()
    for k in ["mixrate", "maxspeed", "max_jump", "action_vx", "action_vy", "action_vrot", "center_x", "center_y"]:
        out[k] = r.f()
    for k in ["random_agent_start", "has_useful_vel_info", "step_rand_int"]:
        out[k] = r.i()
    out["asset_rand_gen"] = (r.i(), r.s())
    for k in ["main_width", "main_height", "out_of_bounds_object"]:
        out[k] = r.i()
    for k in ["unit", "view_dim", "x_off", "y_off", "visibility", "min_visibility"]:
        out[k] = r.f()
    w, h = r.i(), r.i()
    ng = r.i()
    out["grid_w"], out["grid_h"] = w, h
    out["grid"] = list(struct.unpack_from(f"<{ng}i", r.d, r.o))
    r.o += 4 * ng
    out["tail_offset"] = r.o
    return out
