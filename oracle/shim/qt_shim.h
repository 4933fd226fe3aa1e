// TEST INFRASTRUCTURE — not part of the product path.
//
// Minimal stand-in for the slice of Qt 5 (QtGui: QPainter/QImage/QColor/QRect/QRectF) that the
// reference's game sources use, so that /root/reference/procgen/src/**/*.cpp compile UNMODIFIED,
// in place, into oracle/_ref/libenv_ref.so (game logic = the reference's own code; pixels = the
// CPU restatement of Qt's non-antialiased raster rules in qt_raster.cpp).
//
// Third-party dependency being restated: Qt 5.13.2 qtbase (static build for wheels,
// procgen-build/procgen_build/build_qt.py:60) / qt=5.12.5 (environment.yml:9). Not present in
// /root/reference or in this image. PARITY UNPINNED against Qt 5.13.2; the rules are pinned
// against a real Qt 6.6.3 raster engine (oracle/qt6_backend.cpp) — see DESIGN.md.
//
// API surface follows the reference call sites: game.cpp:77-91, basic-abstract-game.cpp:799-1066,
// games/{chaser,jumper,ninja,plunder,starpilot}.cpp custom draws, resources.cpp:19-28,
// assetgen.cpp.
#pragma once
// (real Qt headers pull these in transitively; the reference relies on that)
#include <cstdint>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <map>
#include <set>
#include <list>
#include <algorithm>
#include <functional>
#include <memory>
#include <string>
#include <vector>

// The stand-in classes are renamed at the preprocessor level so their mangled symbols can never
// collide with (or interpose) the real Qt symbols when the Qt 6 backend links libQt6Gui.
#define QString pgshim_QString
#define QPointF pgshim_QPointF
#define QRect pgshim_QRect
#define QRectF pgshim_QRectF
#define QColor pgshim_QColor
#define QBrush pgshim_QBrush
#define QPen pgshim_QPen
#define QImage pgshim_QImage
#define QPainter pgshim_QPainter

typedef unsigned char uchar;
typedef double qreal;

class QString {
  public:
    QString() {}
    QString(const char *s) : str(s) {}
    std::string str;
};

class QPointF {
  public:
    QPointF() : xp(0), yp(0) {}
    QPointF(qreal x, qreal y) : xp(x), yp(y) {}
    qreal x() const { return xp; }
    qreal y() const { return yp; }
    qreal xp, yp;
};

class QRect {
  public:
    QRect() : x1(0), y1(0), w(0), h(0) {}
    QRect(int x, int y, int width, int height) : x1(x), y1(y), w(width), h(height) {}
    int x() const { return x1; }
    int y() const { return y1; }
    int width() const { return w; }
    int height() const { return h; }
    int x1, y1, w, h;
};

class QRectF {
  public:
    QRectF() : xp(0), yp(0), w(0), h(0) {}
    QRectF(qreal x, qreal y, qreal width, qreal height) : xp(x), yp(y), w(width), h(height) {}
    QRectF(const QRect &r) : xp(r.x()), yp(r.y()), w(r.width()), h(r.height()) {}
    qreal x() const { return xp; }
    qreal y() const { return yp; }
    qreal width() const { return w; }
    qreal height() const { return h; }
    QPointF center() const { return QPointF(xp + w / 2, yp + h / 2); }
    qreal xp, yp, w, h;
};

class QColor {
  public:
    QColor() : r(0), g(0), b(0), a(255), valid(false) {}
    QColor(int r_, int g_, int b_, int a_ = 255) : r(r_), g(g_), b(b_), a(a_), valid(true) {}
    void setAlpha(int alpha) { a = alpha; }
    int red() const { return r; }
    int green() const { return g; }
    int blue() const { return b; }
    int alpha() const { return a; }
    int r, g, b, a;
    bool valid;
};

namespace Qt {
enum PenStyle { NoPen = 0, SolidLine = 1 };
}

class QBrush {
  public:
    QBrush() : on(false) {}
    QBrush(const QColor &c) : color(c), on(true) {}
    QColor color;
    bool on;
};

class QPen {
  public:
    QPen() : width(1), on(true) {}
    QPen(Qt::PenStyle s) : width(1), on(s != Qt::NoPen) {}
    QPen(const QColor &c, qreal w = 1) : color(c), width(w), on(true) {}
    QColor color;
    qreal width;
    bool on;
};

struct QtShimImageData;

class QImage {
  public:
    enum Format {
        Format_Invalid = 0,
        Format_RGB32 = 4,
        Format_ARGB32 = 5,
        Format_ARGB32_Premultiplied = 6,
    };
    QImage();
    QImage(int width, int height, Format format);
    QImage(uchar *data, int width, int height, int bytesPerLine, Format format);
    explicit QImage(const QString &fileName);
    QImage(const QImage &o) = default;
    QImage &operator=(const QImage &o) = default;
    ~QImage();

    QImage convertToFormat(Format f) const;
    QImage mirrored(bool horizontal, bool vertical) const;
    int width() const { return w; }
    int height() const { return h; }
    Format format() const { return fmt; }
    uint32_t *pixels() const { return ext ? ext : (store ? store->data() : nullptr); }
    int stride_px() const { return stride; }

    int w = 0, h = 0, stride = 0;
    Format fmt = Format_Invalid;
    uint32_t *ext = nullptr;  // caller-owned pixels (render target)
    std::shared_ptr<std::vector<uint32_t>> store;
    std::shared_ptr<void> backend;  // backend-private handle (e.g. a real Qt6 QImage)
};

class QPainter {
  public:
    enum RenderHint { Antialiasing = 1, SmoothPixmapTransform = 4 };
    enum CompositionMode { CompositionMode_SourceOver = 0, CompositionMode_Source = 3 };
    explicit QPainter(QImage *device);
    ~QPainter();
    void setRenderHint(RenderHint hint, bool on = true);
    void fillRect(const QRectF &r, const QColor &c);
    void fillRect(const QRect &r, const QColor &c);
    void drawImage(const QRectF &target, const QImage &image);
    void save();
    void restore();
    void setOpacity(qreal o);
    void translate(qreal dx, qreal dy);
    void rotate(qreal degrees);
    void setBrush(const QBrush &b);
    void setBrush(const QColor &c) { setBrush(QBrush(c)); }
    void setPen(const QPen &p);
    void setPen(const QColor &c) { setPen(QPen(c)); }
    void setPen(Qt::PenStyle s) { setPen(QPen(s)); }
    void drawEllipse(const QRectF &r);
    void drawEllipse(const QRect &r) { drawEllipse(QRectF(r)); }
    void drawLine(int x1, int y1, int x2, int y2);  // the only four-number overload Qt has: callers truncate
    void setCompositionMode(CompositionMode m);

    struct State;
    State *d;
};
