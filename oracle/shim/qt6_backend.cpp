// TEST INFRASTRUCTURE — alternative implementation of the qt_shim.h classes that forwards every
// QPainter call to a REAL Qt raster engine: the header-less Qt 6.6.3 shared libraries bundled with
// Nsight Compute. Linking this file instead of qt_raster.cpp gives oracle/_ref/libenv_ref_qt6.so =
// the reference's own draw code on a real Qt raster, the pin for the restated rules (the
// reference's Qt 5.13.2 itself is not obtainable offline; see DESIGN.md "pixel parity").
//
// Qt is called through hand-declared C symbols bound to the mangled C++ names (no Qt headers
// exist here). Object layouts used: QImage 32 B, QPainter 8 B, QBrush 8 B, QPen 8 B (all opaque),
// QRectF = 4 doubles {x,y,w,h}, QPointF = 2 doubles, QLineF = 4 doubles,
// QColor = {int32 spec = 1 (Rgb); uint16 a, r, g, b, pad} with 8-bit values * 0x101.
// PNG decoding + premultiplication stay in this file's QImage (same code as qt_raster.cpp).
#define QT_SHIM_QT6_BACKEND 1
#include "qt_raster.cpp"  // QImage + asset pack reader + helpers; its QPainter is compiled out below

extern "C" {
struct RQImage { alignas(8) unsigned char opaque[32]; };
struct RQPainter { void *d; };
struct RQBrush { void *d; };
struct RQPen { void *d; };
struct RQRectF { double x, y, w, h; };
struct RQPointF { double x, y; };
struct RQLine { int x1, y1, x2, y2; };  // QLine: two QPoints of ints
struct RQColor { int32_t spec; uint16_t a, r, g, b, pad; };

void rq_image_ctor(RQImage *self, unsigned char *data, int w, int h, long long bpl, int format, void (*cleanup)(void *), void *info)
    asm("_ZN6QImageC1EPhiixNS_6FormatEPFvPvES2_");
void rq_image_dtor(RQImage *self) asm("_ZN6QImageD1Ev");
void rq_painter_ctor(RQPainter *self, void *paint_device) asm("_ZN8QPainterC1EP12QPaintDevice");
void rq_painter_dtor(RQPainter *self) asm("_ZN8QPainterD1Ev");
void rq_painter_draw_image(RQPainter *self, const RQRectF *target, const RQImage *img, const RQRectF *src, int flags)
    asm("_ZN8QPainter9drawImageERK6QRectFRK6QImageS2_6QFlagsIN2Qt19ImageConversionFlagEE");
void rq_painter_fill_rect(RQPainter *self, const RQRectF *r, const RQColor *c) asm("_ZN8QPainter8fillRectERK6QRectFRK6QColor");
void rq_painter_set_opacity(RQPainter *self, double o) asm("_ZN8QPainter10setOpacityEd");
void rq_painter_rotate(RQPainter *self, double a) asm("_ZN8QPainter6rotateEd");
void rq_painter_translate(RQPainter *self, const RQPointF *p) asm("_ZN8QPainter9translateERK7QPointF");
void rq_painter_save(RQPainter *self) asm("_ZN8QPainter4saveEv");
void rq_painter_restore(RQPainter *self) asm("_ZN8QPainter7restoreEv");
void rq_painter_set_render_hint(RQPainter *self, int hint, bool on) asm("_ZN8QPainter13setRenderHintENS_10RenderHintEb");
void rq_painter_draw_ellipse(RQPainter *self, const RQRectF *r) asm("_ZN8QPainter11drawEllipseERK6QRectF");
void rq_painter_draw_lines(RQPainter *self, const RQLine *lines, int n) asm("_ZN8QPainter9drawLinesEPK5QLinei");
void rq_painter_set_brush(RQPainter *self, const RQBrush *b) asm("_ZN8QPainter8setBrushERK6QBrush");
void rq_painter_set_pen(RQPainter *self, const RQPen *p) asm("_ZN8QPainter6setPenERK4QPen");
void rq_painter_set_pen_style(RQPainter *self, int style) asm("_ZN8QPainter6setPenEN2Qt8PenStyleE");
void rq_painter_set_comp(RQPainter *self, int mode) asm("_ZN8QPainter18setCompositionModeENS_15CompositionModeE");
void rq_brush_ctor_color(RQBrush *self, const RQColor *c, int style) asm("_ZN6QBrushC1ERK6QColorN2Qt10BrushStyleE");
void rq_brush_dtor(RQBrush *self) asm("_ZN6QBrushD1Ev");
void rq_pen_ctor(RQPen *self, const RQBrush *b, double width, int style, int cap, int join)
    asm("_ZN4QPenC1ERK6QBrushdN2Qt8PenStyleENS3_11PenCapStyleENS3_12PenJoinStyleE");
void rq_pen_dtor(RQPen *self) asm("_ZN4QPenD1Ev");
}

static RQColor to_rq(const QColor &c) {
    RQColor q;
    q.spec = 1;
    q.a = (uint16_t)(c.alpha() * 0x101);
    q.r = (uint16_t)(c.red() * 0x101);
    q.g = (uint16_t)(c.green() * 0x101);
    q.b = (uint16_t)(c.blue() * 0x101);
    q.pad = 0;
    return q;
}

struct QPainter::State {
    RQImage dev;
    RQPainter p;
};

QPainter::QPainter(QImage *device) : d(new State) {
    rq_image_ctor(&d->dev, (unsigned char *)device->pixels(), device->width(), device->height(),
                  (long long)device->stride_px() * 4, (int)device->format(), nullptr, nullptr);
    rq_painter_ctor(&d->p, &d->dev);
}
QPainter::~QPainter() {
    rq_painter_dtor(&d->p);
    rq_image_dtor(&d->dev);
    delete d;
}
void QPainter::setRenderHint(RenderHint hint, bool on) { rq_painter_set_render_hint(&d->p, (int)hint, on); }
void QPainter::save() { rq_painter_save(&d->p); }
void QPainter::restore() { rq_painter_restore(&d->p); }
void QPainter::setOpacity(qreal o) { rq_painter_set_opacity(&d->p, o); }
void QPainter::translate(qreal dx, qreal dy) {
    RQPointF pt{dx, dy};
    rq_painter_translate(&d->p, &pt);
}
void QPainter::rotate(qreal a) { rq_painter_rotate(&d->p, a); }
void QPainter::setCompositionMode(CompositionMode m) { rq_painter_set_comp(&d->p, (int)m); }
void QPainter::fillRect(const QRect &r, const QColor &c) { fillRect(QRectF(r), c); }
void QPainter::fillRect(const QRectF &r, const QColor &c) {
    RQRectF rr{r.x(), r.y(), r.width(), r.height()};
    RQColor cc = to_rq(c);
    rq_painter_fill_rect(&d->p, &rr, &cc);
}
void QPainter::drawImage(const QRectF &target, const QImage &image) {
    RQImage src;
    rq_image_ctor(&src, (unsigned char *)image.pixels(), image.width(), image.height(), (long long)image.stride_px() * 4,
                  (int)image.format(), nullptr, nullptr);
    RQRectF t{target.x(), target.y(), target.width(), target.height()};
    RQRectF s{0, 0, (double)image.width(), (double)image.height()};  // inline overload: full source rect
    rq_painter_draw_image(&d->p, &t, &src, &s, 0);
    rq_image_dtor(&src);
}
void QPainter::setBrush(const QBrush &b) {
    RQColor c = to_rq(b.color);
    RQBrush rb;
    rq_brush_ctor_color(&rb, &c, b.on ? 1 /*SolidPattern*/ : 0 /*NoBrush*/);
    rq_painter_set_brush(&d->p, &rb);
    rq_brush_dtor(&rb);
}
void QPainter::setPen(const QPen &p) {
    if (!p.on) {
        rq_painter_set_pen_style(&d->p, 0);
        return;
    }
    RQColor c = to_rq(p.color);
    RQBrush rb;
    rq_brush_ctor_color(&rb, &c, 1);
    RQPen rp;
    rq_pen_ctor(&rp, &rb, p.width, 1 /*SolidLine*/, 0x10 /*SquareCap*/, 0x40 /*BevelJoin*/);
    rq_painter_set_pen(&d->p, &rp);
    rq_pen_dtor(&rp);
    rq_brush_dtor(&rb);
}
void QPainter::drawEllipse(const QRectF &r) {
    RQRectF rr{r.x(), r.y(), r.width(), r.height()};
    rq_painter_draw_ellipse(&d->p, &rr);
}
void QPainter::drawLine(int x1, int y1, int x2, int y2) {
    RQLine l{x1, y1, x2, y2};
    rq_painter_draw_lines(&d->p, &l, 1);
}
