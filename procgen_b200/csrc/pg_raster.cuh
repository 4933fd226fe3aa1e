// Software rasteriser: 64x64 RGB observation of one env, composited per pixel (gather) from an
// ordered list of "blits" that restate what the reference asks Qt's raster engine to draw
// (game.cpp:77-91 -> basic-abstract-game.cpp:799-1066).  No framebuffer read-modify-write: every
// output pixel walks the (short, culled) list of blits that can touch it, in draw order, blends in
// registers and is written once as packed RGB.
//
// Raster rules (Qt raster engine, non-antialiased, restated; see oracle/shim/qt_raster.cpp for
// the CPU twin and DESIGN.md for how they are pinned against real Qt 6.6.3):
//   F fillRect -> [qRound(x), qRound(x+w)) x [qRound(y), qRound(y+h))
//   S scaled drawImage -> nearest neighbour, 16.16 fixed point, target snapped to ints (switch)
//   B src-over with BYTE_MUL; O opacity int(o*256) -> (io*255)>>8
//
// Phases (render kernel: one CTA per env; `tid`/`nthreads` are explicit so the same code runs in
// the host debug harness with nthreads = 1; a barrier separates consecutive phases):
//   prepare_camera   logic thread  prepare_for_drawing -> env header (runs in the logic kernel)
//   frame_begin      all threads   thread 0: window + background/overlay blits; threads i<nx / j<ny:
//                                  geometry + pixel span of grid column i / row j
//   frame_build      all threads   warp 0: entities -> blits, culled and compacted in draw order with
//                                  warp ballots; other warps: one blit per visible grid cell and the
//                                  pixel-column/row -> cell lookup tables (fp64 math happens here,
//                                  once per sprite instead of once per pixel)
//   shade_pixel      all threads   the gather
#pragma once
#include "pg_engine.cuh"

namespace pg {

enum BlitKind : uint8_t { BLIT_NONE = 0, BLIT_IMAGE = 1, BLIT_SOLID = 2 };

struct Blit {
    uint8_t x1, y1, w, h;    // device pixels [x1,x1+w) x [y1,y1+h) after clip + Qt's edge guards;
                             // w == 0 <=> nothing to draw (first word = one 32-bit load in the shader)
    uint8_t kind;
    uint8_t mirror;
    uint16_t opacity;        // Qt intOpacity, 256 = opaque path
    int32_t ix, iy;          // 16.16 source step per device pixel
    uint32_t basex, srcy;    // 16.16 source coordinate at (x1, y1)
    uint32_t src;            // IMAGE: texel offset of the sprite in the atlas; SOLID: 0xFFRRGGBB
    uint16_t sw, sh;
};
static_assert(sizeof(Blit) == 32, "Blit is 32 B");

constexpr int MAX_BG_BLITS = 8;
constexpr int MAX_OVERLAY_BLITS = 8;

// camera of one frame = what prepare_for_drawing leaves in the env header
struct Camera {
    float unit, view_dim, x_off, y_off;
};

template <int MAX_CELLS_1D, int MAX_ENT_BLITS>
struct FrameT {
    static constexpr int kMaxCells1D = MAX_CELLS_1D;
    static constexpr int kMaxEntBlits = MAX_ENT_BLITS;   // VISIBLE entity blits (after culling)
    Camera cam;
    int32_t low_x, low_y, nx, ny;   // visible grid window: cells [low_x, low_x+nx) x [low_y, low_y+ny)
    int32_t n_bg, n_ent, n_ent_below, n_overlay;  // n_ent_below = entities with render_z == -1
    int32_t snap;
    int32_t pad;
    // geometry shared by all cells of a column / row (the cell rect is separable)
    double col_x[MAX_CELLS_1D];     // QRectF.x of column i
    double row_y[MAX_CELLS_1D];     // QRectF.y of row j
    double cell_w;                  // QRectF.width == height
    uint8_t col_p1[MAX_CELLS_1D], col_p2[MAX_CELLS_1D];  // device pixel span [p1,p2) of column i
    uint8_t row_p1[MAX_CELLS_1D], row_p2[MAX_CELLS_1D];
    static constexpr int kEntWords = (MAX_ENT_BLITS + 63) / 64;
    uint64_t ent_rowmask[RES_H][kEntWords];  // bit i: visible entity blit i touches this pixel row
    uint64_t ent_colmask[RES_W][kEntWords];  //        ... this pixel column
    uint8_t col_lo[RES_W], col_hi[RES_W];   // window-relative cell columns covering pixel column
    uint8_t row_lo[RES_H], row_hi[RES_H];
    Blit bg[MAX_BG_BLITS];
    Blit overlay[MAX_OVERLAY_BLITS];
    Blit ents[MAX_ENT_BLITS];
    Blit cells[MAX_CELLS_1D * MAX_CELLS_1D];  // [ci * ny + cj], x outer / y inner = draw order
};

// ---- rule S: un-rotated scaled image (qt_scale_image_32bit)
PG_HD void blit_clear(Blit &b) {
    b.x1 = b.y1 = b.w = b.h = 0;
    b.kind = BLIT_NONE;
}

PG_HD void make_image_blit(Blit &b, double tx, double ty, double tw, double th, SpriteDesc sd, bool mirror, int int_opacity, bool snap) {
    blit_clear(b);
    const int sw = sd.w, sh = sd.h;
    if (sw <= 0 || sh <= 0)
        return;
    if (snap) {
        double x = pg_qround(tx);
        double y = pg_qround(ty);
        double w = pg_qround(tx + tw - x);
        double h = pg_qround(ty + th - y);
        tx = x;
        ty = y;
        tw = w;
        th = h;
    }
    if (!(tw > 0) || !(th > 0))
        return;
    // Qt 6.6.3 qt_scale_image_32bit: step and start both come from the source/target ratio in double
    const double sx = (double)sw / tw;
    const double sy = (double)sh / th;
    const int ix = (int)(65536.0 * sx);
    const int iy = (int)(65536.0 * sy);
    int tx1 = pg_qround(tx), ty1 = pg_qround(ty);
    int tx2 = pg_qround(tx + tw), ty2 = pg_qround(ty + th);
    if (tx1 < 0) tx1 = 0;
    if (ty1 < 0) ty1 = 0;
    if (tx2 > RES_W) tx2 = RES_W;
    if (ty2 > RES_H) ty2 = RES_H;
    if (tx2 <= tx1 || ty2 <= ty1)
        return;
    int h = ty2 - ty1;
    int w = tx2 - tx1;
    const int dstx = (int)pg_dceil((tx1 + 0.5 - tx) * sx * 65536) - 1;
    const int dsty = (int)pg_dceil((ty1 + 0.5 - ty) * sy * 65536) - 1;
    const uint32_t basex = (uint32_t)dstx;
    const uint32_t srcy = (uint32_t)dsty;
    int yend = ((int)(srcy + (uint32_t)iy * (uint32_t)(h - 1))) >> 16;
    if (yend < 0 || yend >= sh)
        --h;
    int xend = ((int)(basex + (uint32_t)ix * (uint32_t)(w - 1))) >> 16;
    if (xend < 0 || xend >= sw)
        --w;
    if (w <= 0 || h <= 0)
        return;
    b.x1 = (uint8_t)tx1;
    b.y1 = (uint8_t)ty1;
    b.w = (uint8_t)w;
    b.h = (uint8_t)h;
    b.kind = BLIT_IMAGE;
    b.mirror = mirror ? 1 : 0;
    b.opacity = (uint16_t)int_opacity;
    b.ix = ix;
    b.iy = iy;
    b.basex = basex;
    b.srcy = srcy;
    b.src = sd.off;
    b.sw = (uint16_t)sw;
    b.sh = (uint16_t)sh;
}

// ---- rule F: opaque fillRect
PG_HD void make_solid_blit(Blit &b, double x, double y, double w, double h, uint32_t rgb) {
    int x1 = pg_qround(x), y1 = pg_qround(y);
    int x2 = pg_qround(x + w), y2 = pg_qround(y + h);
    if (x2 < x1) { int t = x1; x1 = x2; x2 = t; }
    if (y2 < y1) { int t = y1; y1 = y2; y2 = t; }
    if (x1 < 0) x1 = 0;
    if (y1 < 0) y1 = 0;
    if (x2 > RES_W) x2 = RES_W;
    if (y2 > RES_H) y2 = RES_H;
    blit_clear(b);
    if (x2 <= x1 || y2 <= y1)
        return;
    b.x1 = (uint8_t)x1;
    b.y1 = (uint8_t)y1;
    b.w = (uint8_t)(x2 - x1);
    b.h = (uint8_t)(y2 - y1);
    b.kind = BLIT_SOLID;
    b.mirror = 0;
    b.opacity = 256;
    b.src = 0xff000000u | rgb;
    b.ix = b.iy = 0;
    b.basex = b.srcy = 0;
    b.sw = b.sh = 0;
}

// ---- rules B + O
PG_HD uint32_t blend_px(uint32_t dst, uint32_t src, int int_opacity) {
    if (int_opacity == 256) {
        if (src >= 0xff000000u)
            return src;
        if (src != 0)
            return src + pg_byte_mul(dst, (~src) >> 24);
        return dst;
    }
    if (src != 0) {
        uint32_t s = pg_byte_mul(src, (uint32_t)((int_opacity * 255) >> 8));
        return s + pg_byte_mul(dst, (~s) >> 24);
    }
    return dst;
}

PG_HD uint32_t apply_blit(const Blit &b, int px, int py, uint32_t dst, const uint32_t *atlas) {
    const uint32_t box = *reinterpret_cast<const uint32_t *>(&b);  // x1 | y1<<8 | w<<16 | h<<24
    const uint32_t dx = (uint32_t)px - (box & 0xffu);
    const uint32_t dy = (uint32_t)py - ((box >> 8) & 0xffu);
    if (dx >= ((box >> 16) & 0xffu) || dy >= (box >> 24))
        return dst;
    if (b.kind == BLIT_SOLID)
        return b.src;
    uint32_t sx = (b.basex + (uint32_t)b.ix * dx) >> 16;
    uint32_t sy = (b.srcy + (uint32_t)b.iy * dy) >> 16;
    if (b.mirror)
        sx = b.sw - 1 - sx;
    uint32_t texel = atlas[b.src + sy * b.sw + sx];
    return blend_px(dst, texel, b.opacity);
}

template <class G, class Frame>
struct Raster {
    using E = Engine<G>;

    static PG_HD Camera camera_of(const EnvHdr &h) {
        Camera cam;
        cam.unit = h.unit;
        cam.view_dim = h.view_dim;
        cam.x_off = h.x_off;
        cam.y_off = h.y_off;
        return cam;
    }
    // basic-abstract-game.cpp:799-801 — float math, then widened to QRectF doubles
    static PG_HD void screen_rect(const Camera &f, float x, float y, float dx, float dy, float render_eps, double *r) {
        r[0] = (double)((x - render_eps) * f.unit - f.x_off);
        r[1] = (double)((f.view_dim - y - render_eps) * f.unit + f.y_off);
        r[2] = (double)((dx + 2 * render_eps) * f.unit);
        r[3] = (double)((dy + 2 * render_eps) * f.unit);
    }
    // basic-abstract-game.cpp:803-805
    static PG_HD void abs_rect(const Camera &f, float x, float y, float dx, float dy, double *r) {
        r[0] = (double)(x * f.unit);
        r[1] = (double)(y * f.unit);
        r[2] = (double)(dx * f.unit);
        r[3] = (double)(dy * f.unit);
    }
    // qt-utils.h:12-19
    static PG_HD void adjust_rect(double *r, const double *adj) {
        double x = r[0] + r[2] * adj[0];
        double y = r[1] + r[3] * adj[1];
        double w = r[2] * adj[2];
        double h = r[3] * adj[3];
        r[0] = x; r[1] = y; r[2] = w; r[3] = h;
    }
    // basic-abstract-game.cpp:811-817
    static PG_HD void object_rect(const Camera &f, const Entity &o, double *r) {
        if (o.use_abs_coords) {
            abs_rect(f, f.view_dim * (o.x - o.rx), f.view_dim * (o.y + o.ry), 2 * f.view_dim * o.rx, 2 * f.view_dim * o.ry, r);
            return;
        }
        screen_rect(f, o.x - o.rx, o.y + o.ry, 2 * o.rx, 2 * o.ry, 0, r);
    }

    // draw_image (basic-abstract-game.cpp:877-913) for the un-rotated, un-tiled case
    static PG_HD void make_sprite_blit(Ctx &c, const Frame &f, Blit &b, double *rect, float rotation, bool is_reflected, int base_type, int theme, float alpha) {
        blit_clear(b);
        int img_type = G::image_for_type(c, base_type);
        if (img_type < 0)
            return;
        if (c.h->options.use_monochrome_assets || img_type >= USE_ASSET_THRESHOLD) {
            if (img_type == SPACE)
                return;
            if (!G::make_grid_obj_blit(c, f, b, rect, img_type, theme))
                c.h->err |= ERR_UNSUPPORTED;
            return;
        }
        if (theme < 0 || theme >= MAX_IMAGE_THEMES) {
            c.h->err |= ERR_FASSERT;
            return;
        }
        int masked_theme = (c.h->options.restrict_themes && !G::should_preserve_type_themes(c, img_type)) ? 0 : theme;
        double adj[4];
        if (G::get_adjusted_image_rect(c, img_type, adj))
            adjust_rect(rect, adj);
        SpriteDesc sd = c.assets->sprites[img_type + masked_theme * MAX_ASSETS];
        if (sd.w == 0) {
            c.h->err |= ERR_UNSUPPORTED;  // reference would synthesise an asset (assetgen.cpp)
            return;
        }
        int io = 256;
        if (alpha != 1)
            io = (int)((double)alpha * 256);
        if (rotation == 0) {
            make_image_blit(b, rect[0], rect[1], rect[2], rect[3], sd, is_reflected, io, f.snap != 0);
        } else {
            c.h->err |= ERR_UNSUPPORTED;  // rotated sprites: not built yet
        }
    }

    // prepare_for_drawing (basic-abstract-game.cpp:819-838). Writes the camera into the env
    // header (those fields are part of the serialized state, :1202-1220). Logic thread.
    static PG_HD void prepare_camera(Ctx &c) {
        EnvHdr &h = *c.h;
        h.center_x = (float)(h.main_width * .5);
        h.center_y = (float)(h.main_height * .5);
        if (h.options.center_agent) {
            G::choose_center(c, h.center_x, h.center_y);
        } else {
            h.visibility = (float)(h.main_width > h.main_height ? h.main_width : h.main_height);
            if (h.visibility < h.min_visibility)
                h.visibility = h.min_visibility;
        }
        float raw_unit = 64 / h.visibility;
        h.unit = (float)((double)raw_unit * ((double)64.0f / 64.0));
        h.view_dim = (float)(64.0 / (double)raw_unit);
        h.x_off = h.unit * (h.center_x - h.view_dim / 2);
        h.y_off = h.unit * (h.center_y - h.view_dim / 2);
    }

    // visible grid window (basic-abstract-game.cpp:926-938); pure function of the header
    static PG_HD void grid_window(const EnvHdr &h, int &low_x, int &low_y, int &nx, int &ny) {
        int high_x, high_y;
        if (h.options.center_agent) {
            float margin = (float)((double)h.visibility / 2.0 + 1);
            low_x = (int)(h.center_x - margin);
            high_x = (int)(h.center_x + margin);
            low_y = (int)(h.center_y - margin);
            high_y = (int)(h.center_y + margin);
        } else {
            low_x = 0;
            high_x = h.main_width - 1;
            low_y = 0;
            high_y = h.main_height - 1;
        }
        nx = high_x - low_x + 1;
        ny = high_y - low_y + 1;
    }

    static PG_HD void span_of(double t, double tw, bool snap, int limit, uint8_t &p1, uint8_t &p2) {
        if (snap) {
            double x = pg_qround(t);
            tw = pg_qround(t + tw - x);
            t = x;
        }
        int a = pg_qround(t), b2 = pg_qround(t + tw);
        if (a < 0) a = 0;
        if (b2 > limit) b2 = limit;
        if (b2 < a) b2 = a;
        p1 = (uint8_t)a;
        p2 = (uint8_t)b2;
    }

    // ---- phase B
    static PG_HD void frame_begin(Ctx &c, Frame &f, bool snap, int tid, int nthreads) {
        EnvHdr &h = *c.h;
        const Camera cam = camera_of(h);
        int low_x, low_y, nx, ny;
        grid_window(h, low_x, low_y, nx, ny);
        bool overflow = false;
        if (nx > Frame::kMaxCells1D) { nx = Frame::kMaxCells1D; overflow = true; }
        if (ny > Frame::kMaxCells1D) { ny = Frame::kMaxCells1D; overflow = true; }
        if (tid == 0) {
            f.cam = cam;
            f.snap = snap ? 1 : 0;
            f.low_x = low_x;
            f.low_y = low_y;
            f.nx = nx;
            f.ny = ny;
            f.n_overlay = 0;
            f.n_bg = 0;
            f.n_ent = 0;
            f.n_ent_below = 0;
            if (overflow)
                h.err |= ERR_BLIT_OVERFLOW;
            if (h.options.use_backgrounds)
                G::make_background_blits(c, f);
            G::make_overlay_blits(c, f);
        }
        // columns by threads 0.., rows by threads from the top end so they land on other lanes
        for (int i = tid; i < nx; i += nthreads) {
            double r[4];
            screen_rect(cam, (float)(low_x + i), (float)(low_y + 1), 1, 1, RENDER_EPS, r);
            f.col_x[i] = r[0];
            if (i == 0)
                f.cell_w = r[2];
            span_of(r[0], r[2], snap, RES_W, f.col_p1[i], f.col_p2[i]);
        }
        for (int jj = tid; jj < ny; jj += nthreads) {
            int j = ny - 1 - jj;
            double r[4];
            screen_rect(cam, (float)low_x, (float)(low_y + j + 1), 1, 1, RENDER_EPS, r);
            f.row_y[j] = r[1];
            span_of(r[1], r[3], snap, RES_H, f.row_p1[j], f.row_p2[j]);
        }
    }

    static PG_HD void cell_lookup(const uint8_t *p1, const uint8_t *p2, int n, int px, uint8_t &lo, uint8_t &hi) {
        int l = 255, hgh = 0;
        for (int i = 0; i < n; i++) {
            if (px >= p1[i] && px < p2[i]) {
                if (l == 255) {
                    l = i;
                    hgh = i;
                } else {
                    if (i < l) l = i;
                    if (i > hgh) hgh = i;
                }
            }
        }
        lo = (uint8_t)l;
        hi = (uint8_t)hgh;
    }

    // One blit for entity `ei`, or kind NONE when it is not drawn / off screen.
    static PG_HD void entity_blit(Ctx &c, const Frame &f, int ei, Blit &b) {
        blit_clear(b);
        if (!G::should_draw_entity(c, ei))
            return;
        const Entity &o = c.ents[ei];
        double r[4];
        object_rect(f.cam, o, r);
        float tile_ratio = G::get_tile_aspect_ratio(c, ei);
        if (tile_ratio != 0) {
            c.h->err |= ERR_UNSUPPORTED;  // tiled entities: not built yet
            return;
        }
        make_sprite_blit(c, f, b, r, o.rotation, o.is_reflected != 0, o.image_type, o.image_theme, o.alpha);
    }

    // Entities -> visible blits in draw order (draw_entities z=-1 / 0 / 1, basic-abstract-game.cpp:
    // 1059-1066), culled. `lane`/`gsize`: the cooperating group (a full warp on the device, 1 in the
    // host harness). Stable compaction uses the group's ballot.
    static PG_HD void build_entity_blits(Ctx &c, Frame &f, int lane, int gsize) {
        const int n = c.h->n_ents;
        int count = 0;
        int below = 0;
        for (int z = -1; z <= 1; z++) {
            for (int base = 0; base < n; base += gsize) {
                const int ei = base + lane;
                Blit b;
                blit_clear(b);
                if (ei < n && c.ents[ei].render_z == z)
                    entity_blit(c, f, ei, b);
                const bool keep = b.kind != BLIT_NONE;
#if defined(__CUDA_ARCH__)
                const unsigned mask = __ballot_sync(0xffffffffu, keep);
                const int pos = count + __popc(mask & ((1u << lane) - 1u));
                const int total = __popc(mask);
#else
                const int pos = count;
                const int total = keep ? 1 : 0;
#endif
                if (keep) {
                    if (pos < Frame::kMaxEntBlits)
                        f.ents[pos] = b;
                    else
                        c.h->err |= ERR_BLIT_OVERFLOW;
                }
                count += total;
            }
            if (z == -1)
                below = count;
        }
        if (count > Frame::kMaxEntBlits)
            count = Frame::kMaxEntBlits;
        if (below > count)
            below = count;
        if (lane == 0) {
            f.n_ent = count;
            f.n_ent_below = below;
        }
    }

    // ---- phase C. Threads [0, ent_group) build the entity list; the rest build cells + lookups.
    static PG_HD void frame_build(Ctx &c, Frame &f, int tid, int nthreads, int ent_group) {
        if (tid < ent_group) {
            build_entity_blits(c, f, tid, ent_group);
            if (nthreads > ent_group)
                return;
        }
        const int wtid = (nthreads > ent_group) ? tid - ent_group : tid;
        const int wn = (nthreads > ent_group) ? nthreads - ent_group : nthreads;
        for (int px = wtid; px < RES_W + RES_H; px += wn) {
            if (px < RES_W)
                cell_lookup(f.col_p1, f.col_p2, f.nx, px, f.col_lo[px], f.col_hi[px]);
            else
                cell_lookup(f.row_p1, f.row_p2, f.ny, px - RES_W, f.row_lo[px - RES_W], f.row_hi[px - RES_W]);
        }
        const int ncells = f.nx * f.ny;
        for (int k = wtid; k < ncells; k += wn) {
            int ci = k / f.ny, cj = k - ci * f.ny;
            Blit &b = f.cells[k];
            blit_clear(b);
            if (f.col_p1[ci] >= f.col_p2[ci] || f.row_p1[cj] >= f.row_p2[cj])
                continue;  // entirely off screen
            int type = E::get_obj(c, f.low_x + ci, f.low_y + cj);
            if (type == INVALID_OBJ)
                continue;
            int theme = G::theme_for_grid_obj(c, type);
            double r[4] = {f.col_x[ci], f.row_y[cj], f.cell_w, f.cell_w};
            make_sprite_blit(c, f, b, r, 0, false, type, theme, 1.0f);
        }
    }

    // ---- phase C2: per pixel row / column bit masks of the visible entity blits, so a pixel
    // only walks the blits whose box really contains it (rowmask & colmask).
    static PG_HD void frame_masks(Frame &f, int tid, int nthreads) {
        const int n = f.n_ent;
        for (int t = tid; t < RES_H + RES_W; t += nthreads) {
            const bool is_row = t < RES_H;
            const uint32_t q = (uint32_t)(is_row ? t : t - RES_H);
            uint64_t m[Frame::kEntWords];
            for (int w = 0; w < Frame::kEntWords; w++) m[w] = 0;
            for (int i = 0; i < n; i++) {
                const uint32_t box = *reinterpret_cast<const uint32_t *>(&f.ents[i]);
                const uint32_t d = is_row ? q - ((box >> 8) & 0xffu) : q - (box & 0xffu);
                const uint32_t ext = is_row ? (box >> 24) : ((box >> 16) & 0xffu);
                if (d < ext)
                    m[i >> 6] |= (uint64_t)1 << (i & 63);
            }
            for (int w = 0; w < Frame::kEntWords; w++) {
                if (is_row)
                    f.ent_rowmask[q][w] = m[w];
                else
                    f.ent_colmask[q][w] = m[w];
            }
        }
    }

    static PG_HD int ctz64(uint64_t m) {
#if defined(__CUDA_ARCH__)
        return __ffsll((long long)m) - 1;
#else
        return __builtin_ctzll(m);
#endif
    }

    // ---- phase D: the gather. Returns 0xFFRRGGBB (Format_RGB32).
    static PG_HD uint32_t shade_pixel(const Frame &f, int px, int py, const uint32_t *atlas) {
        uint32_t dst = 0xff000000u;  // fillRect(rect, black), basic-abstract-game.cpp:980
        for (int i = 0; i < f.n_bg; i++) dst = apply_blit(f.bg[i], px, py, dst, atlas);
        uint64_t above[Frame::kEntWords];
        const int nb = f.n_ent_below;
        for (int w = 0; w < Frame::kEntWords; w++) {
            uint64_t m = f.ent_rowmask[py][w] & f.ent_colmask[px][w];
            // entities with render_z == -1 go under the grid
            const int lo = nb - w * 64;
            const uint64_t below_bits = lo <= 0 ? 0 : (lo >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << lo) - 1));
            uint64_t mb = m & below_bits;
            above[w] = m & ~below_bits;
            while (mb) {
                const int i = ctz64(mb);
                mb &= mb - 1;
                dst = apply_blit(f.ents[w * 64 + i], px, py, dst, atlas);
            }
        }
        const int clo = f.col_lo[px], chi = f.col_hi[px];
        const int rlo = f.row_lo[py], rhi = f.row_hi[py];
        if (clo != 255 && rlo != 255) {
            for (int ci = clo; ci <= chi; ci++)
                for (int cj = rlo; cj <= rhi; cj++)
                    dst = apply_blit(f.cells[ci * f.ny + cj], px, py, dst, atlas);
        }
        for (int w = 0; w < Frame::kEntWords; w++) {
            uint64_t ma = above[w];
            while (ma) {
                const int i = ctz64(ma);
                ma &= ma - 1;
                dst = apply_blit(f.ents[w * 64 + i], px, py, dst, atlas);
            }
        }
        for (int i = 0; i < f.n_overlay; i++) dst = apply_blit(f.overlay[i], px, py, dst, atlas);
        return dst;
    }
};

// ---- default draw hooks that need Frame (kept out of Defaults<G> to avoid a circular include)
template <class G>
struct DrawDefaults {
    // draw_background's single scaled bg image (basic-abstract-game.cpp:986-1006)
    template <class Frame>
    static PG_HD void make_background_blits(Ctx &c, Frame &f) {
        EnvHdr &h = *c.h;
        double main_rect[4];
        Raster<G, Frame>::screen_rect(f.cam, 0, (float)h.main_height, (float)h.main_width, (float)h.main_height, 0, main_rect);
        SpriteDesc bg = c.assets->backgrounds[h.background_index];
        if (h.bg_tile_ratio < 0) {
            h.err |= ERR_UNSUPPORTED;  // vertical tiling (fruitbot): not built yet
            return;
        }
        float bgw = bg.w;
        float bgh = bg.h;
        float bg_ar = bgw / bgh;
        float world_ar = (float)(h.main_width * 1.0 / h.main_height);
        float extra_w = bg_ar - world_ar;
        float offset_x = h.bg_pct_x * extra_w;
        double adj[4] = {(double)(-offset_x), 0, (double)(bg_ar / world_ar), 1};
        Raster<G, Frame>::adjust_rect(main_rect, adj);
        make_image_blit(f.bg[0], main_rect[0], main_rect[1], main_rect[2], main_rect[3], bg, false, 256, f.snap != 0);
        f.n_bg = 1;
    }
    template <class Frame>
    static PG_HD void make_overlay_blits(Ctx &c, Frame &f) {}
    // draw_grid_obj for types >= 100 (chaser overrides); false = nothing known to draw
    template <class Frame>
    static PG_HD bool make_grid_obj_blit(Ctx &c, const Frame &f, Blit &b, double *rect, int type, int theme) { return false; }
};

}  // namespace pg
