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
// Phases (one CTA per env; `tid`/`nthreads` are explicit so the same code runs in the host debug
// harness with nthreads = 1):
//   frame_setup      one thread   camera (prepare_for_drawing), background blits, z-sorted entity
//                                 order, per-pixel-column/row -> grid-cell lookup
//   frame_build      all threads  one blit per visible grid cell / drawn entity (fp64 math here,
//                                 once per sprite instead of once per pixel)
//   shade_pixel      all threads  the gather
#pragma once
#include "pg_engine.cuh"

namespace pg {

enum BlitKind : uint8_t { BLIT_NONE = 0, BLIT_IMAGE = 1, BLIT_SOLID = 2 };

struct Blit {
    uint8_t x1, y1, x2, y2;  // device pixels [x1,x2) x [y1,y2) after clip + Qt's edge guards
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

template <int MAX_CELLS_1D, int MAX_ENT_BLITS>
struct FrameT {
    static constexpr int kMaxCells1D = MAX_CELLS_1D;
    static constexpr int kMaxEntBlits = MAX_ENT_BLITS;
    // camera
    float unit, view_dim, x_off, y_off;
    int32_t low_x, low_y, nx, ny;   // visible grid window: cells [low_x, low_x+nx) x [low_y, low_y+ny)
    int32_t n_bg, n_ent, n_ent_below, n_overlay;  // n_ent_below = entities with render_z == -1
    int32_t snap;
    int32_t pad;
    // geometry shared by all cells of a column / row (the cell rect is separable)
    double col_x[MAX_CELLS_1D];     // QRectF.x of column i
    double row_y[MAX_CELLS_1D];     // QRectF.y of row j
    double cell_w;                  // QRectF.width == height
    uint8_t col_lo[RES_W], col_hi[RES_W];   // window-relative cell columns covering pixel column
    uint8_t row_lo[RES_H], row_hi[RES_H];
    uint16_t ent_order[MAX_ENT_BLITS];      // entity indices in draw order (z=-1, then 0, then 1)
    Blit bg[MAX_BG_BLITS];
    Blit overlay[MAX_OVERLAY_BLITS];
    Blit ents[MAX_ENT_BLITS];
    Blit cells[MAX_CELLS_1D * MAX_CELLS_1D];  // [ci * ny + cj], x outer / y inner = draw order
};

// ---- rule S: un-rotated scaled image (qt_scale_image_32bit)
PG_HD void make_image_blit(Blit &b, double tx, double ty, double tw, double th, SpriteDesc sd, bool mirror, int int_opacity, bool snap) {
    b.kind = BLIT_NONE;
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
    const double sx = tw / (double)sw;
    const double sy = th / (double)sh;
    const int ix = (int)(65536.0 / sx);
    const int iy = (int)(65536.0 / sy);
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
    const int dstx = (int)pg_dceil((tx1 + 0.5 - tx) * ix) - 1;
    const int dsty = (int)pg_dceil((ty1 + 0.5 - ty) * iy) - 1;
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
    b.x2 = (uint8_t)(tx1 + w);
    b.y2 = (uint8_t)(ty1 + h);
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
    b.kind = BLIT_NONE;
    if (x2 <= x1 || y2 <= y1)
        return;
    b.x1 = (uint8_t)x1;
    b.y1 = (uint8_t)y1;
    b.x2 = (uint8_t)x2;
    b.y2 = (uint8_t)y2;
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
    if (b.kind == BLIT_NONE || px < b.x1 || px >= b.x2 || py < b.y1 || py >= b.y2)
        return dst;
    if (b.kind == BLIT_SOLID)
        return b.src;
    uint32_t sx = (b.basex + (uint32_t)b.ix * (uint32_t)(px - b.x1)) >> 16;
    uint32_t sy = (b.srcy + (uint32_t)b.iy * (uint32_t)(py - b.y1)) >> 16;
    if (b.mirror)
        sx = b.sw - 1 - sx;
    uint32_t texel = atlas[b.src + sy * b.sw + sx];
    return blend_px(dst, texel, b.opacity);
}

template <class G, class Frame>
struct Raster {
    using E = Engine<G>;

    // basic-abstract-game.cpp:799-801 — float math, then widened to QRectF doubles
    static PG_HD void screen_rect(const Frame &f, float x, float y, float dx, float dy, float render_eps, double *r) {
        r[0] = (double)((x - render_eps) * f.unit - f.x_off);
        r[1] = (double)((f.view_dim - y - render_eps) * f.unit + f.y_off);
        r[2] = (double)((dx + 2 * render_eps) * f.unit);
        r[3] = (double)((dy + 2 * render_eps) * f.unit);
    }
    // basic-abstract-game.cpp:803-805
    static PG_HD void abs_rect(const Frame &f, float x, float y, float dx, float dy, double *r) {
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
    static PG_HD void object_rect(Ctx &c, const Frame &f, const Entity &o, double *r) {
        if (o.use_abs_coords) {
            abs_rect(f, f.view_dim * (o.x - o.rx), f.view_dim * (o.y + o.ry), 2 * f.view_dim * o.rx, 2 * f.view_dim * o.ry, r);
            return;
        }
        screen_rect(f, o.x - o.rx, o.y + o.ry, 2 * o.rx, 2 * o.ry, 0, r);
    }

    // draw_image (basic-abstract-game.cpp:877-913) for the un-rotated, un-tiled case
    static PG_HD void make_sprite_blit(Ctx &c, const Frame &f, Blit &b, double *rect, float rotation, bool is_reflected, int base_type, int theme, float alpha) {
        b.kind = BLIT_NONE;
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

    // prepare_for_drawing (basic-abstract-game.cpp:819-838) + draw_background (:979-1007) +
    // the z-ordering of draw_entities (:1052-1066). Runs on one thread.
    static PG_HD void frame_setup(Ctx &c, Frame &f, bool snap) {
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
        f.unit = h.unit;
        f.view_dim = h.view_dim;
        f.x_off = h.x_off;
        f.y_off = h.y_off;
        f.snap = snap ? 1 : 0;
        f.n_overlay = 0;

        // ---- background
        f.n_bg = 0;
        if (h.options.use_backgrounds)
            G::make_background_blits(c, f);

        // ---- visible grid window (basic-abstract-game.cpp:926-938)
        int low_x, high_x, low_y, high_y;
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
        f.low_x = low_x;
        f.low_y = low_y;
        f.nx = high_x - low_x + 1;
        f.ny = high_y - low_y + 1;
        if (f.nx > Frame::kMaxCells1D || f.ny > Frame::kMaxCells1D) {
            h.err |= ERR_BLIT_OVERFLOW;
            if (f.nx > Frame::kMaxCells1D) f.nx = Frame::kMaxCells1D;
            if (f.ny > Frame::kMaxCells1D) f.ny = Frame::kMaxCells1D;
        }
        for (int i = 0; i < RES_W; i++) {
            f.col_lo[i] = 255; f.col_hi[i] = 0;
            f.row_lo[i] = 255; f.row_hi[i] = 0;
        }
        for (int i = 0; i < f.nx; i++) {
            double r[4];
            screen_rect(f, (float)(low_x + i), (float)(low_y + 1), 1, 1, RENDER_EPS, r);
            f.col_x[i] = r[0];
            f.cell_w = r[2];
            double tx = r[0], tw = r[2];
            if (snap) {
                double x = pg_qround(tx);
                tw = pg_qround(tx + tw - x);
                tx = x;
            }
            int a = pg_qround(tx), b2 = pg_qround(tx + tw);
            if (a < 0) a = 0;
            if (b2 > RES_W) b2 = RES_W;
            for (int px = a; px < b2; px++) {
                if (f.col_lo[px] == 255) {
                    f.col_lo[px] = f.col_hi[px] = (uint8_t)i;
                } else {
                    if (i < f.col_lo[px]) f.col_lo[px] = (uint8_t)i;
                    if (i > f.col_hi[px]) f.col_hi[px] = (uint8_t)i;
                }
            }
        }
        for (int j = 0; j < f.ny; j++) {
            double r[4];
            screen_rect(f, (float)low_x, (float)(low_y + j + 1), 1, 1, RENDER_EPS, r);
            f.row_y[j] = r[1];
            double ty = r[1], th = r[3];
            if (snap) {
                double y = pg_qround(ty);
                th = pg_qround(ty + th - y);
                ty = y;
            }
            int a = pg_qround(ty), b2 = pg_qround(ty + th);
            if (a < 0) a = 0;
            if (b2 > RES_H) b2 = RES_H;
            for (int py = a; py < b2; py++) {
                if (f.row_lo[py] == 255) {
                    f.row_lo[py] = f.row_hi[py] = (uint8_t)j;
                } else {
                    if (j < f.row_lo[py]) f.row_lo[py] = (uint8_t)j;
                    if (j > f.row_hi[py]) f.row_hi[py] = (uint8_t)j;
                }
            }
        }

        // ---- entity draw order: stable by render_z in {-1, 0, 1}
        int n = 0;
        for (int z = -1; z <= 1; z++) {
            for (int i = 0; i < h.n_ents; i++) {
                if (c.ents[i].render_z == z) {
                    if (n < Frame::kMaxEntBlits)
                        f.ent_order[n++] = (uint16_t)i;
                    else
                        h.err |= ERR_BLIT_OVERFLOW;
                }
            }
            if (z == -1)
                f.n_ent_below = n;
        }
        f.n_ent = n;
        G::make_overlay_blits(c, f);
    }

    // One blit per visible cell and per entity; independent, so spread over the CTA.
    static PG_HD void frame_build(Ctx &c, Frame &f, int tid, int nthreads) {
        const int ncells = f.nx * f.ny;
        for (int k = tid; k < ncells; k += nthreads) {
            int ci = k / f.ny, cj = k - ci * f.ny;
            Blit &b = f.cells[k];
            b.kind = BLIT_NONE;
            int type = E::get_obj(c, f.low_x + ci, f.low_y + cj);
            if (type == INVALID_OBJ)
                continue;
            int theme = G::theme_for_grid_obj(c, type);
            double r[4] = {f.col_x[ci], f.row_y[cj], f.cell_w, f.cell_w};
            make_sprite_blit(c, f, b, r, 0, false, type, theme, 1.0f);
        }
        for (int k = tid; k < f.n_ent; k += nthreads) {
            int ei = f.ent_order[k];
            Blit &b = f.ents[k];
            b.kind = BLIT_NONE;
            if (!G::should_draw_entity(c, ei))
                continue;
            const Entity &o = c.ents[ei];
            double r[4];
            object_rect(c, f, o, r);
            float tile_ratio = G::get_tile_aspect_ratio(c, ei);
            if (tile_ratio != 0) {
                c.h->err |= ERR_UNSUPPORTED;  // tiled entities: not built yet
                continue;
            }
            make_sprite_blit(c, f, b, r, o.rotation, o.is_reflected != 0, o.image_type, o.image_theme, o.alpha);
        }
    }

    // The gather. Returns 0xFFRRGGBB (Format_RGB32).
    static PG_HD uint32_t shade_pixel(const Frame &f, int px, int py, const uint32_t *atlas) {
        uint32_t dst = 0xff000000u;  // fillRect(rect, black), basic-abstract-game.cpp:980
        for (int i = 0; i < f.n_bg; i++) dst = apply_blit(f.bg[i], px, py, dst, atlas);
        for (int i = 0; i < f.n_ent_below; i++) dst = apply_blit(f.ents[i], px, py, dst, atlas);
        const int clo = f.col_lo[px], chi = f.col_hi[px];
        const int rlo = f.row_lo[py], rhi = f.row_hi[py];
        if (clo != 255 && rlo != 255) {
            for (int ci = clo; ci <= chi; ci++)
                for (int cj = rlo; cj <= rhi; cj++)
                    dst = apply_blit(f.cells[ci * f.ny + cj], px, py, dst, atlas);
        }
        for (int i = f.n_ent_below; i < f.n_ent; i++) dst = apply_blit(f.ents[i], px, py, dst, atlas);
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
        Raster<G, Frame>::screen_rect(f, 0, (float)h.main_height, (float)h.main_width, (float)h.main_height, 0, main_rect);
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
