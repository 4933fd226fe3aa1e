// Host-side asset pipeline: asset pack -> device atlas + per-game metadata tables.
//
// Replaces the reference's QImage loading (resources.cpp:19-28, 813-953) and the lazy
// per-game asset table (basic-abstract-game.cpp:79-123): every (type, theme) sprite a game can
// name is resolved once at init to {atlas offset, w, h}, its aspect ratio and theme count —
// which game LOGIC consumes (basic-abstract-game.cpp:1014-1046) — and pixels are converted with
// Qt's exact premultiply (BYTE_MUL) to ARGB32_Premultiplied / RGB32 words.
// Pure host C++ (no CUDA): shared by the runtime and by the CPU debug harness in tests/.
#pragma once
#include <zlib.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "pg_state.cuh"

namespace pg {
namespace host {

struct PackImage {
    uint32_t w = 0, h = 0;
    uint64_t off = 0;
    uint32_t csize = 0;
};

class AssetPackReader {
  public:
    explicit AssetPackReader(const std::string &path) : path_(path) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f)
            throw std::runtime_error("procgen_b200: cannot open asset pack '" + path + "' (build it with __graft_entry__.build())");
        struct Hdr {
            char magic[8];
            uint32_t version, count;
            uint64_t moff, mlen;
        } hdr;
        if (fread(&hdr, sizeof(hdr), 1, f) != 1 || memcmp(hdr.magic, "PGB2PACK", 8) != 0 || hdr.version != 1) {
            fclose(f);
            throw std::runtime_error("procgen_b200: bad asset pack header in " + path);
        }
        for (uint32_t i = 0; i < hdr.count; i++) {
            struct __attribute__((packed)) Ent {
                char name[112];
                uint32_t w, h;
                uint64_t off;
                uint32_t csize, reserved;
            } e;
            if (fread(&e, sizeof(e), 1, f) != 1)
                break;
            PackImage im;
            im.w = e.w;
            im.h = e.h;
            im.off = e.off;
            im.csize = e.csize;
            index_[std::string(e.name, strnlen(e.name, sizeof(e.name)))] = im;
        }
        // manifest: only the background group lists are needed; tiny hand parser for
        // {"groups": {"name": ["a", "b"], ...}, ...}
        std::string manifest(hdr.mlen, '\0');
        fseek(f, (long)hdr.moff, SEEK_SET);
        if (fread(&manifest[0], 1, hdr.mlen, f) != hdr.mlen) {
            fclose(f);
            throw std::runtime_error("procgen_b200: truncated asset pack " + path);
        }
        fclose(f);
        parse_groups(manifest);
    }

    bool has(const std::string &name) const { return index_.count(name) != 0; }
    const PackImage &info(const std::string &name) const {
        auto it = index_.find(name);
        if (it == index_.end())
            throw std::runtime_error("procgen_b200: asset '" + name + "' not in pack");
        return it->second;
    }
    const std::vector<std::string> &group(const std::string &g) const {
        auto it = groups_.find(g);
        if (it == groups_.end())
            throw std::runtime_error("procgen_b200: background group '" + g + "' not in pack");
        return it->second;
    }
    // straight-alpha RGBA8, row-major
    std::vector<unsigned char> rgba(const std::string &name) const {
        const PackImage &im = info(name);
        std::vector<unsigned char> comp(im.csize), raw((size_t)im.w * im.h * 4);
        FILE *f = fopen(path_.c_str(), "rb");
        if (!f)
            throw std::runtime_error("procgen_b200: cannot reopen asset pack");
        fseek(f, (long)im.off, SEEK_SET);
        size_t got = fread(comp.data(), 1, im.csize, f);
        fclose(f);
        uLongf rawlen = raw.size();
        if (got != im.csize || uncompress(raw.data(), &rawlen, comp.data(), im.csize) != Z_OK || rawlen != raw.size())
            throw std::runtime_error("procgen_b200: corrupt asset '" + name + "'");
        return raw;
    }

  private:
    void parse_groups(const std::string &m) {
        size_t g = m.find("\"groups\"");
        if (g == std::string::npos)
            return;
        size_t pos = m.find('{', g);
        int depth = 0;
        std::string key;
        while (pos < m.size()) {
            char ch = m[pos];
            if (ch == '{') {
                depth++;
                pos++;
            } else if (ch == '}') {
                depth--;
                pos++;
                if (depth == 0)
                    break;
            } else if (ch == '"') {
                size_t e = m.find('"', pos + 1);
                std::string s = m.substr(pos + 1, e - pos - 1);
                pos = e + 1;
                size_t nx = m.find_first_not_of(" \t\n", pos);
                if (nx != std::string::npos && m[nx] == ':') {
                    key = s;
                    groups_[key];
                    pos = nx + 1;
                } else {
                    groups_[key].push_back(s);
                }
            } else {
                pos++;
            }
        }
    }
    std::string path_;
    std::map<std::string, PackImage> index_;
    std::map<std::string, std::vector<std::string>> groups_;
};

inline uint32_t byte_mul8(uint32_t c, uint32_t a) {
    uint32_t t = c * a;
    return (t + (t >> 8) + 0x80) >> 8;
}

// ---------------------------------------------------------------- per-game asset names
// Restated tables of games/*.cpp `asset_for_type` + `load_background_images`
// (+ reserved_asset_for_type, basic-abstract-game.cpp:416-430).
struct GameAssetNames {
    const char *bg_group = nullptr;
    std::map<int, std::vector<std::string>> by_type;
};

inline void add_reserved(GameAssetNames &g) {
    const char *ex[5] = {"misc_assets/explosion1.png", "misc_assets/explosion2.png", "misc_assets/explosion3.png",
                         "misc_assets/explosion4.png", "misc_assets/explosion5.png"};
    for (int i = 0; i < 5; i++)
        if (!g.by_type.count(EXPLOSION + i))
            g.by_type[EXPLOSION + i] = {ex[i]};
    if (!g.by_type.count(TRAIL))
        g.by_type[TRAIL] = {"misc_assets/iconCircle_white.png"};
}

inline std::string lower(std::string s) {
    for (auto &ch : s) ch = (char)tolower((unsigned char)ch);
    return s;
}

GameAssetNames game_asset_names(int game_id);  // defined in pg_asset_tables.h

// ---------------------------------------------------------------- atlas builder
struct AtlasBuilder {
    const AssetPackReader &pack;
    std::vector<uint32_t> texels;                      // device atlas, u32 per texel
    std::map<std::string, SpriteDesc> sprite_cache;    // premultiplied sprites
    std::map<std::string, SpriteDesc> bg_cache;        // RGB32 backgrounds
    std::vector<SpriteDesc> tile_sprites;              // rows of the pre-scaled tile table (one per distinct sprite)
    std::map<uint32_t, int> slot_of_offset;

    explicit AtlasBuilder(const AssetPackReader &p) : pack(p) {}

    SpriteDesc add(const std::string &name, bool premultiplied) {
        auto &cache = premultiplied ? sprite_cache : bg_cache;
        auto it = cache.find(name);
        if (it != cache.end())
            return it->second;
        const PackImage &im = pack.info(name);
        if (im.w > 65535 || im.h > 65535)
            throw std::runtime_error("asset too large: " + name);
        std::vector<unsigned char> raw = pack.rgba(name);
        SpriteDesc d;
        d.off = (uint32_t)texels.size();
        d.w = (uint16_t)im.w;
        d.h = (uint16_t)im.h;
        texels.resize(texels.size() + (size_t)im.w * im.h);
        uint32_t *out = &texels[d.off];
        for (size_t i = 0; i < (size_t)im.w * im.h; i++) {
            uint32_t r = raw[i * 4 + 0], g = raw[i * 4 + 1], b = raw[i * 4 + 2], a = raw[i * 4 + 3];
            if (premultiplied)  // QImage::convertToFormat(Format_ARGB32_Premultiplied), resources.cpp:814
                out[i] = (a << 24) | (byte_mul8(r, a) << 16) | (byte_mul8(g, a) << 8) | byte_mul8(b, a);
            else                // Format_RGB32: colour kept, alpha forced opaque, resources.cpp:946
                out[i] = 0xff000000u | (r << 16) | (g << 8) | b;
        }
        cache[name] = d;
        return d;
    }

    void build_game(int game_id, GameAssets &ga) {
        memset(&ga, 0, sizeof(ga));
        for (auto &sl : ga.sprite_slot) sl = -1;
        GameAssetNames names = game_asset_names(game_id);
        add_reserved(names);
        for (auto &kv : names.by_type) {
            int type = kv.first;
            const auto &list = kv.second;
            if (type < 0 || type >= MAX_ASSETS)
                throw std::runtime_error("bad asset table");
            // asset_num_themes is the full list length even when it exceeds the MAX_IMAGE_THEMES
            // slots the reference can address (dodgeball lists 11 enemies and draws from the first 7)
            ga.num_themes[type] = (int32_t)list.size();
            for (size_t theme = 0; theme < list.size() && theme < (size_t)MAX_IMAGE_THEMES; theme++) {
                SpriteDesc d = add(list[theme], true);
                int idx = type + (int)theme * MAX_ASSETS;
                ga.sprites[idx] = d;
                auto it = slot_of_offset.find(d.off);
                if (it == slot_of_offset.end()) {
                    it = slot_of_offset.emplace(d.off, (int)tile_sprites.size()).first;
                    tile_sprites.push_back(d);
                }
                if (it->second < 32767)
                    ga.sprite_slot[idx] = (int16_t)it->second;
                // basic-abstract-game.cpp:114: width() * 1.0 / height(), stored to a float
                ga.aspect[idx] = (float)(d.w * 1.0 / d.h);
            }
        }
        // types with no names get a generated asset with one theme in the reference
        // (basic-abstract-game.cpp:100-110); logic only ever needs num_themes = 1 for them.
        for (int t = 0; t < MAX_ASSETS; t++)
            if (ga.num_themes[t] == 0) {
                ga.num_themes[t] = 1;
                ga.aspect[t] = 1.0f;
            }
        const auto &bgs = pack.group(names.bg_group);
        if (bgs.size() > (size_t)MAX_BACKGROUNDS)
            throw std::runtime_error("too many backgrounds");
        ga.num_backgrounds = (int32_t)bgs.size();
        for (size_t i = 0; i < bgs.size(); i++) ga.backgrounds[i] = add(bgs[i], false);
    }
};

}  // namespace host
}  // namespace pg
