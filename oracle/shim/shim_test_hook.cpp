// TEST INFRASTRUCTURE — C entry point that issues, through whichever QPainter backend this
// library was linked with (restatement or real Qt 6), exactly the call sequence of
// BasicAbstractGame::draw_image (basic-abstract-game.cpp:877-913) for one sprite. Used by
// tests/test_oracle.py to pin the raster rules on synthetic rect / angle / opacity sweeps.
#include "qt_shim.h"

extern "C" __attribute__((visibility("default"))) void shim_test_draw_image(
    uint32_t *dst, int w, int h, const uint32_t *src, int sw, int sh, int src_premultiplied, double x, double y, double tw,
    double th, double rotation_deg, double opacity, int use_rotation) {
    QImage dev((uchar *)dst, w, h, w * 4, QImage::Format_RGB32);
    QImage sprite((uchar *)src, sw, sh, sw * 4, src_premultiplied ? QImage::Format_ARGB32_Premultiplied : QImage::Format_RGB32);
    QPainter p(&dev);
    if (opacity != 1) {
        p.save();
        p.setOpacity(opacity);
    }
    if (!use_rotation) {
        p.drawImage(QRectF(x, y, tw, th), sprite);
    } else {
        p.save();
        p.translate(x + tw / 2, y + th / 2);
        p.rotate(rotation_deg);
        p.drawImage(QRectF(-tw / 2, -th / 2, tw, th), sprite);
        p.restore();
    }
    if (opacity != 1)
        p.restore();
}

extern "C" __attribute__((visibility("default"))) void shim_test_fill_rect(uint32_t *dst, int w, int h, double x, double y,
                                                                            double rw, double rh, int r, int g, int b) {
    QImage dev((uchar *)dst, w, h, w * 4, QImage::Format_RGB32);
    QPainter p(&dev);
    p.fillRect(QRectF(x, y, rw, rh), QColor(r, g, b));
}

// drawEllipse / drawLine exactly as jumper.cpp:137-169 issues them: set_pen_brush_color (pen and
// brush of one colour, integer pen width) or brush-only with Qt::NoPen.
extern "C" __attribute__((visibility("default"))) void shim_test_draw_ellipse(uint32_t *dst, int w, int h, double x, double y, double rw,
                                                                               double rh, int r, int g, int b, int a, int pen_width) {
    QImage dev((uchar *)dst, w, h, w * 4, QImage::Format_RGB32);
    QPainter p(&dev);
    p.setBrush(QBrush(QColor(r, g, b, a)));
    if (pen_width < 0)
        p.setPen(Qt::NoPen);
    else
        p.setPen(QPen(QColor(r, g, b, a), pen_width));
    p.drawEllipse(QRectF(x, y, rw, rh));
}

extern "C" __attribute__((visibility("default"))) void shim_test_draw_line(uint32_t *dst, int w, int h, int x1, int y1, int x2, int y2,
                                                                            int r, int g, int b, int pen_width) {
    QImage dev((uchar *)dst, w, h, w * 4, QImage::Format_RGB32);
    QPainter p(&dev);
    p.setBrush(QBrush(QColor(r, g, b)));
    p.setPen(QPen(QColor(r, g, b), pen_width));
    p.drawLine(x1, y1, x2, y2);
}
