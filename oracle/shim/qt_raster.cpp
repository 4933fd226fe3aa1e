// TEST INFRASTRUCTURE — CPU restatement of the Qt raster-engine rules the reference's draw code
// relies on (non-antialiased QPainter on a Format_RGB32 QImage; game.cpp:77-91).
//
// Restated from qtbase (QRasterPaintEngine::drawImage / fillRect, qt_scale_image_32bit,
// qblendfunctions, qdrawhelper fetchTransformed) — third-party, absent from /root/reference,
// pinned by the reference at Qt 5.13.2 (procgen-build/procgen_build/build_qt.py:60).  Every rule
// below is cross-checked bit-for-bit against a real Qt 6.6.3 raster engine by
// oracle/shim/qt6_backend.cpp + tests/test_oracle.py.  PARITY UNPINNED against Qt 5.13.2 itself.
//
// Rules (SURVEY §8a R1–R4):
//  F  fillRect(QRectF, opaque): pixels [qRound(x), qRound(x+w)) x [qRound(y), qRound(y+h)).
//  S  un-rotated drawImage: nearest neighbour in 16.16 fixed point; target rect optionally
//     snapped to integers first (QT_SHIM_SNAP, default on = Qt 6.6.3 behaviour).
//  B  blend of ARGB32_Premultiplied onto RGB32: dst = src + BYTE_MUL(dst, 255 - src.a).
//  O  setOpacity(o): io = int(o*256); if io != 256: src = BYTE_MUL(src, (io*255)>>8).
//  R  rotated drawImage: pixel covered iff its centre inverse-maps into the rect; texel by
//     16.16 fixed-point stepping of the inverse map.
#include "qt_shim.h"

#include <zlib.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

// ---------------------------------------------------------------- helpers

// Qt 6's qRound (qnumeric.h): half away from zero. Qt 5 rounded negative ties UP
// (int(d - double(int(d-1)) + 0.5) + int(d-1)); the two differ only for values exactly on -n.5.
// The oracle is pinned against Qt 6.6.3, so its form is the one restated.
static inline int qRound(double d) {
    return d >= 0.0 ? int(d + 0.5) : int(d - 0.5);
}

// QT_SHIM_NODRAW=1 turns every draw call into a no-op: the "logic only" CPU baseline row of
// SURVEY §8(d) (an upper bound on the reference's speed with a free rasteriser). Measurement only.
static inline bool nodraw() {
    static const bool v = getenv("QT_SHIM_NODRAW") != nullptr;
    return v;
}

static inline uint32_t BYTE_MUL(uint32_t x, uint32_t a) {
    uint32_t t = (x & 0xff00ff) * a;
    t = (t + ((t >> 8) & 0xff00ff) + 0x800080) >> 8;
    t &= 0xff00ff;
    x = ((x >> 8) & 0xff00ff) * a;
    x = (x + ((x >> 8) & 0xff00ff) + 0x800080);
    x &= 0xff00ff00;
    return x | t;
}

static inline uint32_t qPremultiply(uint32_t x) {
    const uint32_t a = x >> 24;
    uint32_t t = (x & 0xff00ff) * a;
    t = (t + ((t >> 8) & 0xff00ff) + 0x800080) >> 8;
    t &= 0xff00ff;
    x = ((x >> 8) & 0xff) * a;
    x = (x + ((x >> 8) & 0xff) + 0x80);
    x &= 0xff00;
    return x | t | (a << 24);
}

static bool snap_enabled() {
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("QT_SHIM_SNAP");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v == 1;
}

// ---------------------------------------------------------------- asset pack reader

namespace {
struct PackEntry {
    uint32_t w, h;
    uint64_t off;
    uint32_t csize;
};
struct Pack {
    std::string path;
    std::map<std::string, PackEntry> index;
};
std::mutex g_pack_mutex;
std::map<std::string, std::shared_ptr<Pack>> g_packs;

std::shared_ptr<Pack> open_pack(const std::string &path) {
    std::lock_guard<std::mutex> lock(g_pack_mutex);
    auto it = g_packs.find(path);
    if (it != g_packs.end())
        return it->second;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) {
        fprintf(stderr, "qt_shim: cannot open asset pack %s\n", path.c_str());
        return nullptr;
    }
    auto pack = std::make_shared<Pack>();
    pack->path = path;
    struct {
        char magic[8];
        uint32_t version, count;
        uint64_t moff, mlen;
    } hdr;
    if (fread(&hdr, sizeof(hdr), 1, f) != 1 || memcmp(hdr.magic, "PGB2PACK", 8) != 0) {
        fclose(f);
        return nullptr;
    }
    for (uint32_t i = 0; i < hdr.count; i++) {
        struct {
            char name[112];
            uint32_t w, h;
            uint64_t off;
            uint32_t csize, reserved;
        } __attribute__((packed)) e;
        if (fread(&e, sizeof(e), 1, f) != 1)
            break;
        PackEntry pe{e.w, e.h, e.off, e.csize};
        pack->index[std::string(e.name, strnlen(e.name, sizeof(e.name)))] = pe;
    }
    fclose(f);
    g_packs[path] = pack;
    return pack;
}
}  // namespace

// ---------------------------------------------------------------- QImage

QImage::QImage() {}
QImage::~QImage() {}

QImage::QImage(int width, int height, Format format) : w(width), h(height), stride(width), fmt(format) {
    store = std::make_shared<std::vector<uint32_t>>(size_t(width) * height, 0u);
}

QImage::QImage(uchar *data, int width, int height, int bytesPerLine, Format format)
    : w(width), h(height), stride(bytesPerLine / 4), fmt(format), ext((uint32_t *)data) {}

// fileName = "<pack path ending in .pack>:<relpath>" (the oracle driver passes the pack path + ':'
// as resource_root; resources.cpp:20 concatenates).
QImage::QImage(const QString &fileName) {
    const std::string &s = fileName.str;
    size_t pos = s.find(".pack:");
    if (pos == std::string::npos) {
        fprintf(stderr, "qt_shim: resource_root must be '<assets.pack>:' (got %s)\n", s.c_str());
        return;
    }
    auto pack = open_pack(s.substr(0, pos + 5));
    if (!pack)
        return;
    auto it = pack->index.find(s.substr(pos + 6));
    if (it == pack->index.end()) {
        fprintf(stderr, "qt_shim: %s not in pack\n", s.c_str());
        return;
    }
    const PackEntry &e = it->second;
    std::vector<unsigned char> comp(e.csize);
    FILE *f = fopen(pack->path.c_str(), "rb");
    if (!f)
        return;
    fseek(f, (long)e.off, SEEK_SET);
    size_t got = fread(comp.data(), 1, e.csize, f);
    fclose(f);
    if (got != e.csize)
        return;
    std::vector<unsigned char> raw(size_t(e.w) * e.h * 4);
    uLongf rawlen = raw.size();
    if (uncompress(raw.data(), &rawlen, comp.data(), e.csize) != Z_OK || rawlen != raw.size())
        return;
    w = e.w;
    h = e.h;
    stride = w;
    fmt = Format_ARGB32;
    store = std::make_shared<std::vector<uint32_t>>(size_t(w) * h);
    for (size_t i = 0; i < size_t(w) * h; i++) {
        const unsigned char *p = &raw[i * 4];
        (*store)[i] = (uint32_t(p[3]) << 24) | (uint32_t(p[0]) << 16) | (uint32_t(p[1]) << 8) | p[2];
    }
}

QImage QImage::convertToFormat(Format f) const {
    QImage out(w, h, f);
    const uint32_t *src = pixels();
    if (!src)
        return out;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            uint32_t p = src[y * stride + x];
            if (fmt == Format_ARGB32 && f == Format_ARGB32_Premultiplied)
                p = qPremultiply(p);
            else if (f == Format_RGB32)
                p |= 0xff000000u;
            (*out.store)[size_t(y) * w + x] = p;
        }
    }
    return out;
}

QImage QImage::mirrored(bool horizontal, bool vertical) const {
    QImage out(w, h, fmt);
    const uint32_t *src = pixels();
    if (!src)
        return out;
    for (int y = 0; y < h; y++) {
        int sy = vertical ? h - 1 - y : y;
        for (int x = 0; x < w; x++) {
            int sx = horizontal ? w - 1 - x : x;
            (*out.store)[size_t(y) * w + x] = src[sy * stride + sx];
        }
    }
    return out;
}

#ifndef QT_SHIM_QT6_BACKEND
// ---------------------------------------------------------------- QPainter

struct Xform {
    // Qt convention: x' = m11*x + m21*y + dx ; y' = m12*x + m22*y + dy
    double m11 = 1, m12 = 0, m21 = 0, m22 = 1, dx = 0, dy = 0;
    bool rotated = false;
};

struct PState {
    Xform m;
    double opacity = 1.0;
    QBrush brush;
    QPen pen;
    QPainter::CompositionMode comp = QPainter::CompositionMode_SourceOver;
};

struct QPainter::State {
    QImage *dev;
    PState cur;
    std::vector<PState> stack;
    bool antialias = false;
};

QPainter::QPainter(QImage *device) : d(new State) {
    d->dev = device;
}
QPainter::~QPainter() {
    delete d;
}

void QPainter::setRenderHint(RenderHint hint, bool on) {
    if (hint == Antialiasing)
        d->antialias = on;  // only the out-of-scope 512x512 human render asks for this
}
void QPainter::save() {
    d->stack.push_back(d->cur);
}
void QPainter::restore() {
    if (!d->stack.empty()) {
        d->cur = d->stack.back();
        d->stack.pop_back();
    }
}
void QPainter::setOpacity(qreal o) {
    d->cur.opacity = o;
}
void QPainter::setBrush(const QBrush &b) {
    d->cur.brush = b;
}
void QPainter::setPen(const QPen &p) {
    d->cur.pen = p;
}
void QPainter::setCompositionMode(CompositionMode m) {
    d->cur.comp = m;
}

void QPainter::translate(qreal tx, qreal ty) {
    Xform &m = d->cur.m;
    m.dx += tx * m.m11 + ty * m.m21;
    m.dy += ty * m.m22 + tx * m.m12;
}

void QPainter::rotate(qreal a) {
    // QTransform::rotate: exact values at right angles, sin/cos of a*pi/180 otherwise.
    const double deg2rad = 0.017453292519943295769;
    double sina = 0, cosa = 0;
    if (a == 0.)
        return;
    if (a == 90. || a == -270.)
        sina = 1.;
    else if (a == 270. || a == -90.)
        sina = -1.;
    else if (a == 180.)
        cosa = -1.;
    else {
        double b = deg2rad * a;
        sina = sin(b);
        cosa = cos(b);
    }
    Xform &m = d->cur.m;
    double tm11 = cosa * m.m11 + sina * m.m21;
    double tm12 = cosa * m.m12 + sina * m.m22;
    double tm21 = -sina * m.m11 + cosa * m.m21;
    double tm22 = -sina * m.m12 + cosa * m.m22;
    m.m11 = tm11;
    m.m12 = tm12;
    m.m21 = tm21;
    m.m22 = tm22;
    m.rotated = true;
}

static inline void blend_px(uint32_t *dst, uint32_t src, int int_opacity) {
    if (int_opacity == 256) {
        if (src >= 0xff000000u)
            *dst = src;
        else if (src != 0)
            *dst = src + BYTE_MUL(*dst, (~src) >> 24);
    } else {
        if (src != 0) {
            uint32_t s = BYTE_MUL(src, uint32_t((int_opacity * 255) >> 8));
            *dst = s + BYTE_MUL(*dst, (~s) >> 24);
        }
    }
}

void QPainter::fillRect(const QRect &r, const QColor &c) {
    fillRect(QRectF(r), c);
}

void QPainter::fillRect(const QRectF &r0, const QColor &c) {
    if (nodraw())
        return;
    QImage *dev = d->dev;
    const Xform &m = d->cur.m;
    QRectF r(r0.x() + m.dx, r0.y() + m.dy, r0.width(), r0.height());
    int x1 = qRound(r.x()), y1 = qRound(r.y());
    int x2 = qRound(r.x() + r.width()), y2 = qRound(r.y() + r.height());
    if (x2 < x1)
        std::swap(x1, x2);
    if (y2 < y1)
        std::swap(y1, y2);
    x1 = std::max(x1, 0);
    y1 = std::max(y1, 0);
    x2 = std::min(x2, dev->w);
    y2 = std::min(y2, dev->h);
    uint32_t argb = (uint32_t(c.alpha()) << 24) | (uint32_t(c.red()) << 16) | (uint32_t(c.green()) << 8) | uint32_t(c.blue());
    uint32_t pm = qPremultiply(argb);
    int io = int(d->cur.opacity * 256);
    uint32_t *px = dev->pixels();
    for (int y = y1; y < y2; y++)
        for (int x = x1; x < x2; x++) {
            uint32_t *dst = &px[y * dev->stride + x];
            if (d->cur.comp == CompositionMode_Source)
                *dst = pm;
            else
                blend_px(dst, pm, io);
        }
}

// Un-rotated scaled blit (qt_scale_image_32bit semantics).
static void draw_scaled(QImage *dev, const QImage &img, QRectF tr, int int_opacity) {
    const int sw = img.width(), sh = img.height();
    if (sw <= 0 || sh <= 0)
        return;
    if (snap_enabled()) {
        double x = qRound(tr.x());
        double y = qRound(tr.y());
        double w = qRound(tr.x() + tr.width() - x);
        double h = qRound(tr.y() + tr.height() - y);
        tr = QRectF(x, y, w, h);
    }
    if (tr.width() == 0 || tr.height() == 0)
        return;
    int ix, iy, dstx, dsty;
    int tx1, ty1, tx2, ty2, h, w;
    static int variant = -1;
    if (variant < 0) {
        const char *e = getenv("QT_SHIM_SCALE_VARIANT");
        variant = e ? atoi(e) : 1;
    }
    {
        // targetRect.normalized().toRect()
        double nx = tr.x(), ny = tr.y(), nw = tr.width(), nh = tr.height();
        if (nw < 0) {
            nx += nw;
            nw = -nw;
        }
        if (nh < 0) {
            ny += nh;
            nh = -nh;
        }
        tx1 = qRound(nx), ty1 = qRound(ny);
        tx2 = qRound(nx + nw), ty2 = qRound(ny + nh);
    }
    tx1 = std::max(tx1, 0);
    ty1 = std::max(ty1, 0);
    tx2 = std::min(tx2, dev->w);
    ty2 = std::min(ty2, dev->h);
    if (tx2 <= tx1 || ty2 <= ty1)
        return;
    h = ty2 - ty1;
    w = tx2 - tx1;
    if (variant == 0) {
        // Qt <= 5.x as recalled in SURVEY: step from the target/source ratio, start from the step
        const double sx = tr.width() / double(sw);
        const double sy = tr.height() / double(sh);
        ix = int(0x00010000 / sx);
        iy = int(0x00010000 / sy);
        dstx = int(ceil((tx1 + 0.5 - tr.x()) * ix)) - 1;
        dsty = int(ceil((ty1 + 0.5 - tr.y()) * iy)) - 1;
    } else {
        // Qt 6.6.3 (verified): step and start both from the source/target ratio in double
        const double sx = double(sw) / tr.width();
        const double sy = double(sh) / tr.height();
        ix = int(0x00010000 * sx);
        iy = int(0x00010000 * sy);
        if (sx < 0)  // mirrored: measured from the rect's right() = x + w, source from its right end
            dstx = int(floor((tx1 + 0.5 - (tr.x() + tr.width())) * sx * 65536)) + 1 + sw * 65536;
        else
            dstx = int(ceil((tx1 + 0.5 - tr.x()) * sx * 65536)) - 1;
        if (sy < 0)
            dsty = int(floor((ty1 + 0.5 - (tr.y() + tr.height())) * sy * 65536)) + 1 + sh * 65536;
        else
            dsty = int(ceil((ty1 + 0.5 - tr.y()) * sy * 65536)) - 1;
    }
    uint32_t basex = uint32_t(dstx);
    uint32_t srcy = uint32_t(dsty);
    if (int(srcy >> 16) >= sh && iy < 0) {
        srcy += iy;
        --h;
    }
    if (int(basex >> 16) >= sw && ix < 0) {
        basex += ix;
        --w;
    }

    // Qt's guard against a last row/column sampled just outside the source
    int yend = int(srcy + uint32_t(iy) * uint32_t(h - 1)) >> 16;
    if (yend < 0 || yend >= sh)
        --h;
    int xend = int(basex + uint32_t(ix) * uint32_t(w - 1)) >> 16;
    if (xend < 0 || xend >= sw)
        --w;

    const uint32_t *src = img.pixels();
    uint32_t *px = dev->pixels();
    for (int y = 0; y < h; y++) {
        const uint32_t *srow = src + (srcy >> 16) * img.stride_px();
        uint32_t *drow = px + (ty1 + y) * dev->stride + tx1;
        uint32_t srcx = basex;
        for (int x = 0; x < w; x++) {
            blend_px(&drow[x], srow[srcx >> 16], int_opacity);
            srcx += ix;
        }
        srcy += iy;
    }
}

// ---------------------------------------------------------------- transformed (rotated) blits
// Qt raster engine, non-antialiased, non-smooth drawImage under a rotating world matrix
// (QRasterPaintEngine::drawImage). Two code paths, both restated:
//  (T) device bounding box >= 16x16: qt_transform_image (qblendfunctions_p.h): the quad is cut into
//      three trapezoids whose left/right edges are stepped in 16.16 fixed point; texel coordinates
//      are an absolute affine function of the device pixel in 16.16.
//  (R) smaller targets: QRasterizer::rasterizeLine (qrasterizer.cpp) produces spans for the
//      rotated rectangle (edges stepped in 16.16 from pixel-centre intersections), and
//      fetchTransformed (qdrawhelper.cpp) walks each span in 16.16 from the inverse matrix of
//      (translate(1/65536, 1/65536) * image-to-device).
struct TVertex {
    double x, y, u, v;
};

static inline int qFloorI(double v) { return int(floor(v)); }
static inline int qCeilI(double v) { return int(ceil(v)); }

static void transform_image_rasterize(QImage *dev, const QImage &img, const TVertex &topLeft, const TVertex &bottomLeft,
                                      const TVertex &topRight, const TVertex &bottomRight, double topY, double bottomY, int dudx,
                                      int dvdx, int dudy, int dvdy, int u0, int v0, int int_opacity) {
    const int sw = img.width(), sh = img.height();
    long long fromY = std::max((long long)qRound(topY), 0LL);
    long long toY = std::min((long long)qRound(bottomY), (long long)dev->h);
    if (fromY >= toY)
        return;
    double leftSlope = (bottomLeft.x - topLeft.x) / (bottomLeft.y - topLeft.y);
    double rightSlope = (bottomRight.x - topRight.x) / (bottomRight.y - topRight.y);
    long long dx_l = (long long)(leftSlope * 0x10000);
    long long dx_r = (long long)(rightSlope * 0x10000);
    long long x_l = (long long)((topLeft.x + (0.5 + fromY - topLeft.y) * leftSlope + 0.5) * 0x10000);
    long long x_r = (long long)((topRight.x + (0.5 + fromY - topRight.y) * rightSlope + 0.5) * 0x10000);
    const uint32_t *src = img.pixels();
    uint32_t *px = dev->pixels();
    for (long long y = fromY; y < toY; ++y) {
        long long fromX = std::max(x_l >> 16, 0LL);
        long long toX = std::min(x_r >> 16, (long long)dev->w);
        if (fromX < toX) {
            long long u = fromX * dudx + y * dudy + u0;
            long long v = fromX * dvdx + y * dvdy + v0;
            for (long long x = fromX; x < toX; ++x) {
                long long uu = u >> 16, vv = v >> 16;
                // "Because of rounding, we can get source coordinates outside the source image": clamp
                uu = std::min(std::max(uu, 0LL), (long long)sw - 1);
                vv = std::min(std::max(vv, 0LL), (long long)sh - 1);
                blend_px(&px[y * dev->stride + x], src[vv * img.stride_px() + uu], int_opacity);
                u += dudx;
                v += dvdx;
            }
        }
        x_l += dx_l;
        x_r += dx_r;
    }
}

static void draw_transform_image(QImage *dev, const QImage &img, const QRectF &r, const Xform &m, int int_opacity) {
    const int sw = img.width(), sh = img.height();
    TVertex v[4];  // TopLeft, TopRight, BottomRight, BottomLeft
    v[0].u = v[3].u = 0;
    v[0].v = v[1].v = 0;
    v[1].u = v[2].u = sw;
    v[3].v = v[2].v = sh;
    v[0].x = v[3].x = r.x();
    v[0].y = v[1].y = r.y();
    v[1].x = v[2].x = r.x() + r.width();
    v[3].y = v[2].y = r.y() + r.height();
    for (int i = 0; i < 4; i++) {
        double fx = v[i].x, fy = v[i].y;
        v[i].x = m.m11 * fx + m.m21 * fy + m.dx;
        v[i].y = m.m12 * fx + m.m22 * fy + m.dy;
    }
    int topmost = 0;
    for (int i = 1; i < 4; ++i)
        if (v[i].y < v[topmost].y)
            topmost = i;
    switch (topmost) {
    case 1: {
        TVertex t = v[0];
        for (int i = 0; i < 3; ++i) v[i] = v[i + 1];
        v[3] = t;
    } break;
    case 2:
        std::swap(v[0], v[2]);
        std::swap(v[1], v[3]);
        break;
    case 3: {
        TVertex t = v[3];
        for (int i = 3; i > 0; --i) v[i] = v[i - 1];
        v[0] = t;
    } break;
    }
    // if necessary, swap vertex 1 and 3 such that 1 is to the left of 3
    double dx1 = v[1].x - v[0].x, dy1 = v[1].y - v[0].y;
    double dx2 = v[3].x - v[0].x, dy2 = v[3].y - v[0].y;
    if (dx1 * dy2 - dx2 * dy1 > 0)
        std::swap(v[1], v[3]);

    TVertex u = {v[1].x - v[0].x, v[1].y - v[0].y, v[1].u - v[0].u, v[1].v - v[0].v};
    TVertex w = {v[2].x - v[0].x, v[2].y - v[0].y, v[2].u - v[0].u, v[2].v - v[0].v};
    double det = u.x * w.y - u.y * w.x;
    if (det == 0)
        return;
    double invDet = 1.0 / det;
    double m11 = (u.u * w.y - u.y * w.u) * invDet;
    double m12 = (u.x * w.u - u.u * w.x) * invDet;
    double m21 = (u.v * w.y - u.y * w.v) * invDet;
    double m22 = (u.x * w.v - u.v * w.x) * invDet;
    double mdx = v[0].u - m11 * v[0].x - m12 * v[0].y;
    double mdy = v[0].v - m21 * v[0].x - m22 * v[0].y;
    int dudx = int(m11 * 0x10000);
    int dvdx = int(m21 * 0x10000);
    int dudy = int(m12 * 0x10000);
    int dvdy = int(m22 * 0x10000);
    int u0 = qCeilI((0.5 * m11 + 0.5 * m12 + mdx) * 0x10000) - 1;
    int v0 = qCeilI((0.5 * m21 + 0.5 * m22 + mdy) * 0x10000) - 1;
    if (v[1].y < v[3].y) {
        transform_image_rasterize(dev, img, v[0], v[1], v[0], v[3], v[0].y, v[1].y, dudx, dvdx, dudy, dvdy, u0, v0, int_opacity);
        transform_image_rasterize(dev, img, v[1], v[2], v[0], v[3], v[1].y, v[3].y, dudx, dvdx, dudy, dvdy, u0, v0, int_opacity);
        transform_image_rasterize(dev, img, v[1], v[2], v[3], v[2], v[3].y, v[2].y, dudx, dvdx, dudy, dvdy, u0, v0, int_opacity);
    } else {
        transform_image_rasterize(dev, img, v[0], v[1], v[0], v[3], v[0].y, v[3].y, dudx, dvdx, dudy, dvdy, u0, v0, int_opacity);
        transform_image_rasterize(dev, img, v[0], v[1], v[3], v[2], v[3].y, v[1].y, dudx, dvdx, dudy, dvdy, u0, v0, int_opacity);
        transform_image_rasterize(dev, img, v[1], v[2], v[3], v[2], v[1].y, v[2].y, dudx, dvdx, dudy, dvdy, u0, v0, int_opacity);
    }
}

// ---- path (R): spans from QRasterizer::rasterizeLine, texels from fetchTransformed
struct SpanSink {
    QImage *dev;
    const QImage *img;
    int io;
    // inverse of (translate(1/65536,1/65536) * image->device)
    double m11, m12, m21, m22, dx, dy;
    void span(int x, int len, int y) const {
        if (y < 0 || y >= dev->h)
            return;
        if (x < 0) {
            len += x;
            x = 0;
        }
        if (x + len > dev->w)
            len = dev->w - x;
        if (len <= 0)
            return;
        const int sw = img->width(), sh = img->height();
        const double fixed_scale = 65536.0;
        const double cx = x + 0.5, cy = y + 0.5;
        const int fdx = int(m11 * fixed_scale);
        const int fdy = int(m12 * fixed_scale);
        int fx = int((m21 * cy + m11 * cx + dx) * fixed_scale);
        int fy = int((m22 * cy + m12 * cx + dy) * fixed_scale);
        const uint32_t *src = img->pixels();
        uint32_t *px = dev->pixels();
        for (int i = 0; i < len; i++) {
            int tx = std::min(std::max(fx >> 16, 0), sw - 1);
            int ty = std::min(std::max(fy >> 16, 0), sh - 1);
            blend_px(&px[y * dev->stride + x + i], src[ty * img->stride_px() + tx], io);
            fx += fdx;
            fy += fdy;
        }
    }
};

static inline bool q26Dot6Compare(double p1, double p2) { return int((p2 - p1) * 64.) == 0; }
static inline double qSafeDivide(double x, double y) {
    if (y == 0)
        return x > 0 ? 1e20 : -1e20;
    return x / y;
}
static inline int qSafeFloatToQ16Dot16(double x) {
    double tmp = x * 65536.;
    if (tmp > double(INT32_MAX))
        return INT32_MAX;
    if (tmp < -double(INT32_MAX))
        return -INT32_MAX;
    return int(tmp);
}
static inline int FloatToQ16Dot16(double i) { return int(i * 65536.); }
static inline double qBoundD(double lo, double v, double hi) { return std::max(lo, std::min(hi, v)); }

// QRasterizer::rasterizeLine(a, b, width), non-antialiased, clip = device rect
struct ScanLine {
    int x, delta, top, bottom, winding;
};

static void scan_convert_quad(const SpanSink &sink, const double *vx, const double *vy) {
    const int W = sink.dev->w, H = sink.dev->h;
    long X[4], Y[4];
    for (int i = 0; i < 4; i++) {
        X[i] = (long)((vx[i] - 0.5) * 64);
        Y[i] = (long)((vy[i] - 0.5) * 64);
    }
    ScanLine lines[4];
    int n = 0;
    for (int i = 0; i < 4; i++) {
        long ax = X[i], ay = Y[i], bx = X[(i + 1) & 3], by = Y[(i + 1) & 3];
        if (ax == bx && ay == by)
            continue;
        int winding = 1;
        if (ay > by) {
            std::swap(ax, bx);
            std::swap(ay, by);
            winding = -1;
        }
        ax += 32; ay += 32; bx += 32; by += 32;  // COORD_OFFSET
        int iTop = std::max(0, int((ay + 32 - 1) >> 6));
        int iBottom = std::min(H - 1, int((by - 32 - 1) >> 6));
        if (iTop <= iBottom) {
            int aFP = 0x8000 + int(ax * 1024) - 1;
            if (bx == ax) {
                lines[n++] = ScanLine{aFP, 0, iTop, iBottom, winding};
            } else {
                const double slope = (bx - ax) / double(by - ay);
                const int slopeFP = int(slope * 65536.);
                const long long dy = (long long)(iTop << 16) + 0x8000 - (long long)ay * 1024;
                int xFP = aFP + int(((long long)slopeFP * dy) >> 16);
                lines[n++] = ScanLine{xFP, slopeFP, iTop, iBottom, winding};
            }
        }
    }
    if (n == 0)
        return;
    std::stable_sort(lines, lines + n, [](const ScanLine &a, const ScanLine &b) { return a.top < b.top; });
    ScanLine *active[4];
    int na = 0, li = 0;
    for (int y = lines[0].top; y < H; ++y) {
        for (; li < n && lines[li].top == y; ++li) active[na++] = &lines[li];
        if (na == 0 && li >= n)
            break;
        for (int i = 1; i < na; ++i) {
            ScanLine *t = active[i];
            int j = i;
            while (j > 0 && active[j - 1]->x > t->x) {
                active[j] = active[j - 1];
                --j;
            }
            active[j] = t;
        }
        int x = 0, winding = 0;
        int keep = 0;
        ScanLine *next_active[4];
        for (int i = 0; i < na; ++i) {
            ScanLine *node = active[i];
            const int current = node->x >> 16;
            if (winding & 1) {
                int x0 = std::max(x, 0), x1 = std::min(current, W);
                if (x1 > x0)
                    sink.span(x0, x1 - x0, y);
            }
            x = current;
            winding += node->winding;
            if (node->bottom != y) {
                node->x += node->delta;
                next_active[keep++] = node;
            }
        }
        for (int i = 0; i < keep; i++) active[i] = next_active[i];
        na = keep;
    }
}

static void rasterize_line(const SpanSink &sink, double ax, double ay, double bx, double by, double width) {
    const int clipL = 0, clipT = 0, clipR = sink.dev->w - 1, clipB = sink.dev->h - 1;  // inclusive QRect edges
    if ((ax == bx && ay == by) || width == 0)
        return;
    double pax = ax, pay = ay, pbx = bx, pby = by;
    {
        // clip the segment to the device rect grown by the line's half extent
        const double offx = fabs(by - ay) * width * 0.5, offy = fabs(bx - ax) * width * 0.5;
        const double cl = clipL - offx, ct = clipT - offy, cr = (clipR + 1) + offx, cb = (clipB + 1) + offy;
        auto inside = [&](double x, double y) { return !(x < cl || x > cr || y < ct || y > cb); };
        if (!inside(pax, pay) || !inside(pbx, pby)) {
            double t1 = 0, t2 = 1;
            const double o[2] = {pax, pay};
            const double dd[2] = {pbx - pax, pby - pay};
            const double low[2] = {cl, ct};
            const double high[2] = {cr, cb};
            for (int i = 0; i < 2; ++i) {
                if (dd[i] == 0) {
                    if (o[i] <= low[i] || o[i] >= high[i])
                        return;
                    continue;
                }
                const double d_inv = 1 / dd[i];
                double t_low = (low[i] - o[i]) * d_inv;
                double t_high = (high[i] - o[i]) * d_inv;
                if (t_low > t_high)
                    std::swap(t_low, t_high);
                if (t1 < t_low)
                    t1 = t_low;
                if (t2 > t_high)
                    t2 = t_high;
                if (t1 >= t2)
                    return;
            }
            const double npax = pax + (pbx - pax) * t1, npay = pay + (pby - pay) * t1;
            const double npbx = pax + (pbx - pax) * t2, npby = pay + (pby - pay) * t2;
            pax = npax; pay = npay; pbx = npbx; pby = npby;
        }
        const double d0x = ax - bx, d0y = ay - by;
        const double w0 = d0x * d0x + d0y * d0y;
        const double d1x = pax - pbx, d1y = pay - pby;
        const double w = d1x * d1x + d1y * d1y;
        if (w == 0)
            return;
        width *= sqrt(w0 / w);
    }
    if (q26Dot6Compare(pay, pby)) {
        const double x = (pax + pbx) * 0.5f;
        const double dx = fabs(pbx - pax) * 0.5f;
        const double y = pay;
        const double dy = width * dx;
        pax = x;
        pay = y - dy;
        pbx = x;
        pby = y + dy;
        width = 1 / width;
    }
    if (q26Dot6Compare(pax, pbx)) {
        if (pay > pby) {
            std::swap(pax, pbx);
            std::swap(pay, pby);
        }
        const double dy = pby - pay;
        const double halfWidth = 0.5f * width * dy;
        double left = pax - halfWidth;
        double right = pax + halfWidth;
        left = qBoundD(double(clipL), left, double(clipR + 1));
        right = qBoundD(double(clipL), right, double(clipR + 1));
        pay = qBoundD(double(clipT), pay, double(clipB + 1));
        pby = qBoundD(double(clipT), pby, double(clipB + 1));
        if (q26Dot6Compare(left, right) || q26Dot6Compare(pay, pby))
            return;
        int iTop = int(pay + 0.5f);
        int iBottom = pby < 0.5f ? -1 : int(pby - 0.5f);
        int iLeft = int(left + 0.5f);
        int iRight = right < 0.5f ? -1 : int(right - 0.5f);
        int iWidth = iRight - iLeft + 1;
        for (int y = iTop; y <= iBottom; ++y) sink.span(iLeft, iWidth, y);
        return;
    }
    if (pay > pby) {
        std::swap(pax, pbx);
        std::swap(pay, pby);
    }
    double deltax = (pbx - pax) * (0.5f * width), deltay = (pby - pay) * (0.5f * width);
    const double perpx = deltay, perpy = -deltax;
    double topx, topy, leftx, lefty, rightx, righty, bottomx, bottomy;
    if (pax < pbx) {
        topx = pax + perpx; topy = pay + perpy;
        leftx = pax - perpx; lefty = pay - perpy;
        rightx = pbx + perpx; righty = pby + perpy;
        bottomx = pbx - perpx; bottomy = pby - perpy;
    } else {
        topx = pax - perpx; topy = pay - perpy;
        leftx = pbx - perpx; lefty = pby - perpy;
        rightx = pax + perpx; righty = pay + perpy;
        bottomx = pbx + perpx; bottomy = pby + perpy;
    }
    // general case: the four corners go through the aliased scan converter
    const double vx[4] = {topx, rightx, bottomx, leftx};
    const double vy[4] = {topy, righty, bottomy, lefty};
    scan_convert_quad(sink, vx, vy);
}

static void draw_rotated(QImage *dev, const QImage &img, const QRectF &r, const Xform &m, int int_opacity) {
    const int sw = img.width(), sh = img.height();
    if (sw <= 0 || sh <= 0 || r.width() <= 0 || r.height() <= 0)
        return;
    // targetBounds = s->matrix.mapRect(r): bounding box of the four mapped corners
    double minx = 1e300, miny = 1e300, maxx = -1e300, maxy = -1e300;
    for (int c = 0; c < 4; c++) {
        double fx = (c & 1) ? r.x() + r.width() : r.x(), fy = (c & 2) ? r.y() + r.height() : r.y();
        double X = m.m11 * fx + m.m21 * fy + m.dx, Y = m.m12 * fx + m.m22 * fy + m.dy;
        minx = std::min(minx, X);
        maxx = std::max(maxx, X);
        miny = std::min(miny, Y);
        maxy = std::max(maxy, Y);
    }
    if (maxx - minx >= 16 && maxy - miny >= 16) {
        draw_transform_image(dev, img, r, m, int_opacity);
        return;
    }
    // copy = matrix; copy.translate(r.x, r.y); copy.scale(r.w / sw, r.h / sh)
    double c11 = m.m11, c12 = m.m12, c21 = m.m21, c22 = m.m22;
    double cdx = m.dx + r.x() * m.m11 + r.y() * m.m21;
    double cdy = m.dy + r.y() * m.m22 + r.x() * m.m12;
    const double sx = r.width() / double(sw), sy = r.height() / double(sh);
    c11 *= sx;
    c12 *= sx;
    c21 *= sy;
    c22 *= sy;
    // QSpanData::setupMatrix: inv = (translate(1/65536, 1/65536) * copy).inverted()
    const double t = 1.0 / 65536;
    double pdx = t * c11 + t * c21 + cdx;
    double pdy = t * c12 + t * c22 + cdy;
    double det = c11 * c22 - c12 * c21;
    if (det == 0)
        return;
    double dinv = 1.0 / det;
    SpanSink sink;
    sink.dev = dev;
    sink.img = &img;
    sink.io = int_opacity;
    sink.m11 = c22 * dinv;
    sink.m12 = -c12 * dinv;
    sink.m21 = -c21 * dinv;
    sink.m22 = c11 * dinv;
    sink.dx = (c21 * pdy - c22 * pdx) * dinv;
    sink.dy = (c12 * pdx - c11 * pdy) * dinv;
    // Coverage: QRasterizer::rasterizeLine on the segment joining the mid points of the left and
    // right edges of r (what QRasterPaintEngine::drawImage does for shear-free transforms): exact
    // axis-aligned branch, otherwise the aliased scan converter on the four corners.
    double ly = (r.y() + (r.y() + r.height())) * 0.5f;
    double lx = (r.x() + r.x()) * 0.5f;
    double rx = ((r.x() + r.width()) + (r.x() + r.width())) * 0.5f;
    double ax = m.m11 * lx + m.m21 * ly + m.dx, ay = m.m12 * lx + m.m22 * ly + m.dy;
    double bx = m.m11 * rx + m.m21 * ly + m.dx, by = m.m12 * rx + m.m22 * ly + m.dy;
    rasterize_line(sink, ax, ay, bx, by, r.height() / r.width());
}

void QPainter::drawImage(const QRectF &target, const QImage &image) {
    if (nodraw())
        return;
    const Xform &m = d->cur.m;
    int io = int(d->cur.opacity * 256);
    static const bool trace = getenv("QT_SHIM_TRACE") != nullptr;
    if (trace && d->dev->w == 64)
        fprintf(stderr, "drawImage %.17g %.17g %.17g %.17g img %dx%d rot %d dx %.17g dy %.17g io %d\n", target.x(), target.y(), target.width(),
                target.height(), image.width(), image.height(), (int)m.rotated, m.dx, m.dy, io);
    if (!m.rotated) {
        draw_scaled(d->dev, image, QRectF(target.x() + m.dx, target.y() + m.dy, target.width(), target.height()), io);
    } else if (fabs(m.m12) <= 1e-12 && fabs(m.m21) <= 1e-12) {
        // QTransform::type() is fuzzy (qFuzzyIsNull): rotate(+-180) classifies as TxScale and takes the
        // scaling blit with qt_mapRect_non_normalizing; width/height may be negative
        draw_scaled(d->dev, image,
                    QRectF(m.m11 * target.x() + m.dx, m.m22 * target.y() + m.dy, m.m11 * target.width(), m.m22 * target.height()), io);
    } else {
        draw_rotated(d->dev, image, target, m, io);
    }
}

// ---------------------------------------------------------------- ellipse and cosmetic line
// Used by the jumper compass only (jumper.cpp:137-169). Both are integer algorithms in Qt's raster
// engine and are restated here; tests/test_oracle.py sweeps them against the real Qt 6 backend.

// solid-colour span, clipped to the device (QSpanData solid blend: src-over with the premultiplied colour)
static void solid_span(QImage *dev, int x, int len, int y, uint32_t pm, int io) {
    if (y < 0 || y >= dev->h || len <= 0)
        return;
    int x0 = std::max(x, 0), x1 = std::min(x + len, dev->w);
    uint32_t *row = dev->pixels() + (size_t)y * dev->stride;
    for (int xx = x0; xx < x1; xx++) blend_px(&row[xx], pm, io);
}

static uint32_t premul_color(const QColor &c) {
    uint32_t argb = (uint32_t(c.alpha()) << 24) | (uint32_t(c.red()) << 16) | (uint32_t(c.green()) << 8) | uint32_t(c.blue());
    return qPremultiply(argb);
}

struct EllipseCtx {
    QImage *dev;
    int rx, ry, rw, rh;  // integer rect
    bool pen, brush;
    uint32_t pen_pm, brush_pm;
    int io;
};

// drawEllipsePoints (qpaintengine_raster.cpp): the four mirrored outline spans of one octant step
// plus the two fill spans between them
static void ellipse_points(const EllipseCtx &e, int x, int y, int length) {
    if (length == 0)
        return;
    const int midx = e.rx + (e.rw + 1) / 2;
    const int midy = e.ry + (e.rh + 1) / 2;
    x = x + midx;
    y = midy - y;
    int ox[4], ol[4], oy[4];
    ox[0] = midx + (midx - x) - (length - 1) - (e.rw & 0x1);  // top left
    ol[0] = std::min(length, x - ox[0]);
    oy[0] = y;
    ox[1] = x;  // top right
    ol[1] = length;
    oy[1] = y;
    ox[2] = ox[0];  // bottom left
    ol[2] = ol[0];
    oy[2] = midy + (midy - y) - (e.rh & 0x1);
    ox[3] = x;  // bottom right
    ol[3] = length;
    oy[3] = oy[2];
    if (e.brush && ox[0] + ol[0] < ox[1]) {
        int fx0 = ox[0] + ol[0] - 1;
        int fl0 = std::max(0, ox[1] - fx0);
        int fx1 = ox[2] + ol[2] - 1;
        int fl1 = std::max(0, ox[3] - fx1);
        int n = (oy[1] >= oy[3]) ? 1 : 2;
        solid_span(e.dev, fx0, fl0, oy[1], e.brush_pm, e.io);
        if (n == 2)
            solid_span(e.dev, fx1, fl1, oy[3], e.brush_pm, e.io);
    }
    if (e.pen) {
        int n = (oy[1] >= oy[2]) ? 2 : 4;
        for (int i = 0; i < n; i++) solid_span(e.dev, ox[i], ol[i], oy[i], e.pen_pm, e.io);
    }
}

// drawEllipse_midpoint_i (qpaintengine_raster.cpp)
static void ellipse_midpoint(const EllipseCtx &e) {
    const double a = double(e.rw) / 2;
    const double b = double(e.rh) / 2;
    double d = b * b - (a * a * b) + 0.25 * a * a;
    int x = 0;
    int y = (e.rh + 1) / 2;
    int startx = x;
    // region 1
    while (a * a * (2 * y - 1) > 2 * b * b * (x + 1)) {
        if (d < 0) {
            d += b * b * (2 * x + 3);
            ++x;
        } else {
            d += b * b * (2 * x + 3) + a * a * (-2 * y + 2);
            ellipse_points(e, startx, y, x - startx + 1);
            startx = ++x;
            --y;
        }
    }
    ellipse_points(e, startx, y, x - startx + 1);
    // region 2
    d = b * b * (x + 0.5) * (x + 0.5) + a * a * ((y - 1) * (y - 1) - b * b);
    const int miny = e.rh & 0x1;
    while (y > miny) {
        if (d < 0) {
            d += b * b * (2 * x + 2) + a * a * (-2 * y + 3);
            ++x;
        } else {
            d += a * a * (-2 * y + 3);
        }
        --y;
        ellipse_points(e, x, y, 1);
    }
}

// QRasterPaintEngine::drawEllipse, non-antialiased, unrotated, pen <= 1 px or none.
//  * integer-aligned rect (jumper hard mode's compass disc at (55,1,8,8); every QRect ellipse):
//    Qt 6.6.3 runs the midpoint algorithm above — swept against the real library, all rects.
//  * any other rect goes through Qt's generic path code (Bezier flattening + scan conversion + a
//    cosmetic stroke of the outline). The only such calls in scope are jumper's compass disc on the
//    three non-integer rects the 64x64 contract produces (easy mode's agent-centred view, and the
//    whole-world views of center_agent = false in easy and hard mode): their pixel rows were
//    captured once from Qt 6.6.3 (tests/tools/qt6_compass_mask.py) and are replayed here.
//    Anything else is reported, never approximated.
struct EllipseRowSpan { int x1, x2; };
struct CapturedDisc {
    double x, y, w;
    int first_row, n_rows;
    EllipseRowSpan rows[17];
};
static const CapturedDisc kCompassDiscs[] = {
    {46.66666793823242, 1.3333333730697632, 16.0, 1, 17,
     {{52, 58}, {50, 59}, {49, 60}, {48, 61}, {48, 62}, {47, 63}, {47, 63}, {46, 63}, {46, 63}, {46, 63}, {47, 63}, {47, 63}, {47, 62}, {48, 61},
      {49, 60}, {51, 59}, {53, 57}}},
    {53.60000228881836, 0.800000011920929, 9.600000381469727, 0, 11,
     {{58, 59}, {56, 61}, {55, 62}, {54, 63}, {53, 63}, {53, 64}, {53, 64}, {54, 64}, {54, 63}, {55, 62}, {57, 61}}},
    {60.400001525878906, 0.4000000059604645, 3.200000047683716, 0, 4, {{61, 63}, {60, 64}, {60, 64}, {61, 63}}},
};

void QPainter::drawEllipse(const QRectF &rr) {
    if (nodraw())
        return;
    const Xform &m = d->cur.m;
    const double x = rr.x() * m.m11 + m.dx, y = rr.y() * m.m22 + m.dy, w = rr.width() * m.m11, h = rr.height() * m.m22;
    const int io = int(d->cur.opacity * 256);
    const bool integral = x == std::floor(x) && y == std::floor(y) && w == std::floor(w) && h == std::floor(h);
    if (!integral) {
        const bool same_opaque = d->cur.pen.on && d->cur.brush.on && d->cur.pen.color.alpha() == 255 &&
                                 premul_color(d->cur.pen.color) == premul_color(d->cur.brush.color);
        if (same_opaque && d->dev->w == 64 && d->dev->h == 64) {
            for (const CapturedDisc &c : kCompassDiscs) {
                if (x == c.x && y == c.y && w == c.w && h == c.w) {
                    for (int i = 0; i < c.n_rows; i++)
                        solid_span(d->dev, c.rows[i].x1, c.rows[i].x2 - c.rows[i].x1, c.first_row + i, premul_color(d->cur.pen.color), io);
                    return;
                }
            }
        }
        fprintf(stderr, "qt shim: drawEllipse on a non-integer rect (%.17g %.17g %.17g %.17g) is outside the restated subset\n", x, y, w, h);
        abort();
    }
    EllipseCtx e;
    e.dev = d->dev;
    e.rx = int(x);
    e.ry = int(y);
    e.rw = int(x + w) - int(x);
    e.rh = int(y + h) - int(y);
    if (e.rw <= 0 || e.rh <= 0)
        return;
    e.pen = d->cur.pen.on;
    e.brush = d->cur.brush.on;
    e.pen_pm = premul_color(d->cur.pen.color);
    e.brush_pm = premul_color(d->cur.brush.color);
    e.io = io;
    ellipse_midpoint(e);
}

// QCosmeticStroker::drawLine for one isolated line with square caps (the default QPen cap), integer
// end points, non-antialiased, no dashes. 26.6 end points, 16.16 minor-axis stepping.
static inline int fdot16_div(int x, int y) {
    if (std::abs(x) > 0x7fff)
        return int((long long)x * (1 << 16) / y);
    return x * (1 << 16) / y;
}
void QPainter::drawLine(int ix1, int iy1, int ix2, int iy2) {
    if (nodraw())
        return;
    if (!d->cur.pen.on)
        return;
    QImage *dev = d->dev;
    const Xform &m = d->cur.m;
    const uint32_t pm = premul_color(d->cur.pen.color);
    const int io = int(d->cur.opacity * 256);
    const double rx1 = ix1 * m.m11 + m.dx, ry1 = iy1 * m.m22 + m.dy, rx2 = ix2 * m.m11 + m.dx, ry2 = iy2 * m.m22 + m.dy;
    if (rx1 == rx2 && ry1 == ry2) {  // drawPoints: one pixel
        solid_span(dev, int(std::floor(rx1)), 1, int(std::floor(ry1)), pm, io);
        return;
    }
    enum { CapBegin = 1, CapEnd = 2 };
    int caps = CapBegin | CapEnd;
    int x1 = int(rx1 * 64.), x2 = int(rx2 * 64.), y1 = int(ry1 * 64.), y2 = int(ry2 * 64.);
    const int dx = std::abs(x2 - x1), dy = std::abs(y2 - y1);
    if (dx < dy) {
        if (y1 > y2) {
            std::swap(y1, y2);
            std::swap(x1, x2);
        }
        const int xinc = fdot16_div(x2 - x1, y2 - y1);
        int x = x1 * (1 << 10);
        if (caps & CapBegin) {
            y1 -= 32;
            x -= xinc >> 1;
        }
        if (caps & CapEnd)
            y2 += 32;
        int y = (y1 + 32) >> 6;
        const int ys = (y2 + 32) >> 6;
        const int round = (xinc > 0) ? 32 : 0;
        if (y != ys) {
            x += ((y * (1 << 6)) + round - y1) * xinc >> 6;
            do {
                solid_span(dev, x >> 16, 1, y, pm, io);
                x += xinc;
            } while (++y < ys);
        }
    } else {
        if (!dx)
            return;
        if (x1 > x2) {
            std::swap(x1, x2);
            std::swap(y1, y2);
        }
        const int yinc = fdot16_div(y2 - y1, x2 - x1);
        int y = y1 * (1 << 10);
        if (caps & CapBegin) {
            x1 -= 32;
            y -= yinc >> 1;
        }
        if (caps & CapEnd)
            x2 += 32;
        int x = (x1 + 32) >> 6;
        const int xs = (x2 + 32) >> 6;
        const int round = (yinc > 0) ? 32 : 0;
        if (x != xs) {
            y += ((x * (1 << 6)) + round - x1) * yinc >> 6;
            do {
                solid_span(dev, x, 1, y >> 16, pm, io);
                y += yinc;
            } while (++x < xs);
        }
    }
}

#endif  // QT_SHIM_QT6_BACKEND
