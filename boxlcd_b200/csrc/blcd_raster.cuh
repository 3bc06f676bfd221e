// blcd_raster.cuh -- row-oriented LCD rasterizer (host+device inline).
//
// Replaces WorldEnv.lcd_render (boxLCD/world_env.py:460-512): every dynamic body is drawn as a filled ellipse
// (circle fixtures) or filled polygon into a 1-bit canvas of lcd_w x lcd_h, flipped top-bottom, background = 1.
// The reference delegates the fill to Pillow's C rasterizer; the rules restated here (coordinate pipeline, Bresenham
// ellipse, scanline polygon with apex extension) are the ones pinned against the reference in oracle/lcd_oracle.c.
//
// Formulation: instead of drawing shapes into an image, one call computes the ink mask of ONE canvas row for ONE shape;
// a frame row is the OR over bodies.  A row is one RowMask (bit x = pixel x; uint32, or uint64 in the large-scene
// profile), so a frame is lcd_h coalesced stores
// and no atomics are needed.  The body transform is evaluated in fp32 WITHOUT fused multiply-add (Box2D's
// b2Mul(b2Transform, b2Vec2) on x86-64), the metre->pixel scaling in fp64 with truncation toward zero, as the reference.
#pragma once
#include "blcd_scene.h"

namespace BLCD_NS {

#ifdef __CUDA_ARCH__
#define BLCD_FMUL(a, b) __fmul_rn(a, b)
#define BLCD_FADD(a, b) __fadd_rn(a, b)
#define BLCD_FSUB(a, b) __fsub_rn(a, b)
#else
#define BLCD_FMUL(a, b) ((a) * (b))
#define BLCD_FADD(a, b) ((a) + (b))
#define BLCD_FSUB(a, b) ((a) - (b))
#endif

// (int)(v / WIDTH * width) in fp64 (world_env.py:493-505: numpy float64 arithmetic, then PIL's C cast)
BLCD_HD int to_px(double v, double world_w, double lcd_w) { return (int)(v / world_w * lcd_w); }

// inclusive span [x0, x1] clipped to the window [x_off, x_off + w) as a bit mask (bit 0 = column x_off); ends are swapped if
// inverted (Draw.c hline).  Frames wider than one RowMask are rendered as several such column windows.
BLCD_HD RowMask span_mask(int x0, int x1, int w, int x_off = 0) {
  if (x0 > x1) { int t = x0; x0 = x1; x1 = t; }
  x0 -= x_off; x1 -= x_off;
  if (x0 < 0) x0 = 0;
  if (x1 >= w) x1 = w - 1;
  if (x0 > x1) return 0u;
  int n = x1 - x0 + 1;
  RowMask m = n >= kRowBits ? ~(RowMask)0 : (((RowMask)1 << n) - 1u);
  return m << x0;
}

BLCD_HD int round_up_px(float f) { return f >= 0.0f ? (int)floor((double)f + 0.5) : -(int)floor(-(double)f + 0.5); }
BLCD_HD int round_down_px(float f) { return f >= 0.0f ? (int)ceil((double)f - 0.5) : -(int)ceil(-(double)f - 0.5); }

BLCD_HD long long ell_delta(long long a, long long b, long long x, long long y) {
  long long d = a * a * y * y + b * b * x * x - a * a * b * b;
  return d < 0 ? -d : d;
}

// ink of canvas row y for the filled ellipse inscribed in the integer box (x0, y0, x1, y1): Pillow's ellipseNew walk
// of the first quadrant in doubled coordinates; the widest X reached at height |Y| gives the span of rows +-Y.
BLCD_HD RowMask ellipse_row(int x0, int y0, int x1, int y1, int y, int w, int x_off = 0) {
  int a = x1 - x0, b = y1 - y0;
  if (a < 0 || b < 0 || (a == 0 && b == 0)) return 0u;
  int sy = 2 * (y - y0) - b;
  int Y = sy < 0 ? -sy : sy;
  if (Y > b) return 0u;
  int cx = a, cy = b % 2;
  const int ex = a % 2, ey = b;
  while (cy != Y) {  // cy never decreases; the first arrival at cy == Y carries the largest cx of that row
    if (cx == ex && cy == ey) return 0u;
    int nx = cx, ny = cy + 2;
    long long nd = ell_delta(a, b, nx, ny);
    if (cx > 1) {
      long long d1 = ell_delta(a, b, cx - 2, cy + 2);
      if (nd > d1) { nx = cx - 2; ny = cy + 2; nd = d1; }
      long long d2 = ell_delta(a, b, cx - 2, cy);
      if (nd > d2) { nx = cx - 2; ny = cy; }
    }
    cx = nx; cy = ny;
  }
  return span_mask(x0 + (a - cx) / 2, x0 + (a + cx) / 2, w, x_off);
}

struct PolyPx {
  int n;
  int x[BLCD_MAX_VERTS], y[BLCD_MAX_VERTS];
};

BLCD_HD float edge_x_at(int ex0, int ey0, float dx, int y) { return BLCD_FADD(BLCD_FMUL((float)(y - ey0), dx), (float)ex0); }

// ink of canvas row y for the filled polygon with integer vertices P (Pillow polygon_generic restated per row).
// rules: BLCD_RASTER_PIL12 (pinned) or BLCD_RASTER_PIL9.
// ylo_in / yhi_in: the polygon's vertex row range if the caller already has it (ylo_in > yhi_in: compute here)
BLCD_HD RowMask polygon_row(const PolyPx& P, int y, int w, int h, int rules, int ylo_in = 1, int yhi_in = 0, int x_off = 0) {
  const int n = P.n;
  if (n <= 0) return 0u;
  // edge list: (P[i], P[i+1]) for i < n-1, plus the closing edge unless the last vertex repeats the first
  const int ne = (n > 1 && (P.x[n - 1] != P.x[0] || P.y[n - 1] != P.y[0])) ? n : n - 1;
  if (ne <= 0) return 0u;
  int ylo = ylo_in, yhi = yhi_in;
  if (ylo_in > yhi_in) {
    ylo = P.y[0]; yhi = P.y[0];
    for (int i = 0; i < ne; ++i) {
      int j = i + 1 < n ? i + 1 : 0;
      int a = P.y[i], b = P.y[j];
      ylo = a < ylo ? a : ylo; ylo = b < ylo ? b : ylo;
      yhi = a > yhi ? a : yhi; yhi = b > yhi ? b : yhi;
    }
  }
  RowMask mask = 0u;
  const int Ymin = ylo > 0 ? ylo : 0, Ymax = yhi < h ? yhi : h;
  float xx[2 * BLCD_MAX_VERTS];
  int cnt = 0;
  const bool in_range = (y >= Ymin && y <= Ymax);
  for (int i = 0; i < ne; ++i) {
    int j = i + 1 < n ? i + 1 : 0;
    int ex0 = P.x[i], ey0 = P.y[i], ex1 = P.x[j], ey1 = P.y[j];
    if (ey0 == ey1) {
      if (ey0 == y && rules != BLCD_RASTER_PIL9) mask |= span_mask(ex0, ex1, w, x_off);   // the older Pillow skips horizontal edges
      continue;
    }
    if (!in_range) continue;
    int emin = ey0 < ey1 ? ey0 : ey1, emax = ey0 < ey1 ? ey1 : ey0;
    if (emin <= y && y <= emax) {
      float dx = (float)(ex1 - ex0) / (float)(ey1 - ey0);
      float xv = edge_x_at(ex0, ey0, dx, y);
      // insertion into the sorted list (qsort of <= 16 floats in the original)
      int reps = (y == emax && y < Ymax) ? 2 : 1;
      for (int r = 0; r < reps; ++r) {
        int k = cnt++;
        while (k > 0 && xx[k - 1] > xv) { xx[k] = xx[k - 1]; --k; }
        xx[k] = xv;
      }
    }
  }
  if (!in_range) return mask;
  if (rules == BLCD_RASTER_PIL9) {
    for (int k = 1; k < cnt; k += 2) mask |= span_mask(round_up_px(xx[k - 1]), round_down_px(xx[k]), w, x_off);
    return mask;
  }
  bool have_pos = false;
  int pos = 0;
  for (int k = 1; k < cnt; k += 2) {
    int xs = round_up_px(xx[k - 1]), xe = round_down_px(xx[k]);
    if (have_pos) {
      if (xe < pos) continue;
      if (xs < pos) xs = pos;
    }
    if (xe < xs) continue;
    mask |= span_mask(xs, xe, w, x_off);
    pos = xe + 1; have_pos = true;
  }
  // apex extension: two sloped edges of the same direction meeting in an integer vertex on this row.  Integer tests come
  // first so that the (fp32) slopes are only divided out for edge pairs that really meet on this row.
  for (int i = 1; i < ne; ++i) {
    int ij = i + 1 < n ? i + 1 : 0;
    int cx0 = P.x[i], cy0 = P.y[i], cx1 = P.x[ij], cy1 = P.y[ij];
    if (cy0 == cy1 || cx0 == cx1) continue;                 // horizontal, or slope exactly 0
    int cmin = cy0 < cy1 ? cy0 : cy1, cmax = cy0 < cy1 ? cy1 : cy0;
    if (y != cmin && y != cmax) continue;
    int vx = (cy0 == y) ? cx0 : cx1;
    const bool cpos = ((cx1 - cx0) > 0) == ((cy1 - cy0) > 0);  // sign of dx = (x1 - x0) / (y1 - y0)
    for (int k = 0; k < i; ++k) {
      int kj = k + 1 < n ? k + 1 : 0;
      int ox0 = P.x[k], oy0 = P.y[k], ox1 = P.x[kj], oy1 = P.y[kj];
      if (oy0 == oy1 || ox0 == ox1) continue;
      const bool opos = ((ox1 - ox0) > 0) == ((oy1 - oy0) > 0);
      if (cpos != opos) continue;
      int omin = oy0 < oy1 ? oy0 : oy1, omax = oy0 < oy1 ? oy1 : oy0;
      bool top = (y == cmin && y == omin);
      bool bot = (y == cmax && y == omax && y == Ymax);
      if (!top && !bot) continue;
      int ovx = (oy0 == y) ? ox0 : ox1;
      if (vx != ovx) continue;
      float cdx = (float)(cx1 - cx0) / (float)(cy1 - cy0);
      float odx = (float)(ox1 - ox0) / (float)(oy1 - oy0);
      int y2 = (y == Ymax) ? y - 1 : y + 1;
      float a1 = edge_x_at(cx0, cy0, cdx, y2), a2 = edge_x_at(ox0, oy0, odx, y2);
      if ((bot && cpos) || (top && !cpos)) {
        int s = round_up_px(BLCD_FADD(a1 > a2 ? a1 : a2, 1.0f));
        if (s <= vx) mask |= span_mask(s, vx, w, x_off);
      } else {
        int t = round_up_px(a1 < a2 ? a1 : a2) - 1;
        if (t >= vx) mask |= span_mask(vx, t, w, x_off);
      }
    }
  }
  return mask;
}

// ---- the same scanline rules, written for converged warps ----------------------------------------------------------
// polygon_row above walks data-dependent paths (insertion sort, early `continue`s, a division per crossing edge and row).
// The render kernel gives every lane of a warp the same body, so the edge count is uniform: polygon_row_t<NE> unrolls the
// edge loop, keeps the crossings in registers (slot 2i / 2i+1 of edge i, +inf when unused), sorts them with a fixed
// merge-exchange network and takes the slopes from a per-body table (PolySlopes, one division per edge and BODY instead of
// one per edge and ROW).  Same values, same spans: tests/test_hostsim_vs_oracle.py compares it with polygon_row on millions
// of integer polygons, and the golden frames pin it on the GPU.
struct PolySlopes { float dx[BLCD_MAX_VERTS]; };   // dx[i] = slope of edge P[i] -> P[i + 1 (mod n)], unused for horizontal edges

BLCD_HD void polygon_slopes(PolySlopes& S, const PolyPx& P) {
  for (int i = 0; i < BLCD_MAX_VERTS; ++i) {
    const int j = (i + 1 < P.n && i + 1 < BLCD_MAX_VERTS) ? i + 1 : 0;
    const int dy = P.y[j] - P.y[i];
    S.dx[i] = (i < P.n && dy != 0) ? (float)(P.x[j] - P.x[i]) / (float)dy : 0.0f;
  }
}

// exact integer forms of round_up_px / round_down_px for |f| < 2^23 (the fraction f - floor(f) is exact in fp32)
BLCD_HD int round_up_px_i(float f) {
  const float a = f >= 0.0f ? f : -f, fl = floorf(a);
  const int r = (int)fl + ((a - fl) >= 0.5f ? 1 : 0);
  return f >= 0.0f ? r : -r;
}
BLCD_HD int round_down_px_i(float f) {
  const float a = f >= 0.0f ? f : -f, fl = floorf(a);
  const int r = (int)fl + ((a - fl) > 0.5f ? 1 : 0);
  return f >= 0.0f ? r : -r;
}

template <int NE>
BLCD_HD RowMask polygon_row_t(const PolyPx& P, const PolySlopes& S, int y, int w, int h, int rules, int ylo, int yhi, int x_off) {
  const int n = P.n;
  if (n <= 0) return 0u;
  int xl = P.x[0], yl = P.y[0];
#pragma unroll
  for (int i = 1; i < NE; ++i)
    if (i == n - 1) { xl = P.x[i]; yl = P.y[i]; }
  const int ne = (n > 1 && (xl != P.x[0] || yl != P.y[0])) ? n : n - 1;
  if (ne <= 0) return 0u;
  RowMask mask = 0u;
  const int Ymin = ylo > 0 ? ylo : 0, Ymax = yhi < h ? yhi : h;
  const bool in_range = (y >= Ymin && y <= Ymax);
  const float INF = 3.0e38f;
  float e[2 * NE];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < NE; ++i) {
    const bool last = (i + 1 >= n);
    const int ex0 = P.x[i], ey0 = P.y[i];
    const int ex1 = (last || i + 1 >= NE) ? P.x[0] : P.x[i + 1 < NE ? i + 1 : 0];
    const int ey1 = (last || i + 1 >= NE) ? P.y[0] : P.y[i + 1 < NE ? i + 1 : 0];
    const bool valid = i < ne;
    const bool horizontal = ey0 == ey1;
    if (valid && horizontal && ey0 == y && rules != BLCD_RASTER_PIL9) mask |= span_mask(ex0, ex1, w, x_off);
    const int emin = ey0 < ey1 ? ey0 : ey1, emax = ey0 < ey1 ? ey1 : ey0;
    const bool active = valid && !horizontal && in_range && emin <= y && y <= emax;
    const bool twice = active && y == emax && y < Ymax;
    const float xv = edge_x_at(ex0, ey0, S.dx[i], y);
    e[2 * i] = active ? xv : INF;
    e[2 * i + 1] = twice ? xv : INF;
    cnt += (active ? 1 : 0) + (twice ? 1 : 0);
  }
  if (!in_range) return mask;
  // Batcher's merge-exchange network over the 2 NE slots (compile-time index pairs: everything stays in registers)
  constexpr int N = 2 * NE;
#pragma unroll
  for (int p = 1; p < N; p *= 2) {
#pragma unroll
    for (int k = p; k >= 1; k /= 2) {
#pragma unroll
      for (int j = k % p; j + k < N; j += 2 * k) {
#pragma unroll
        for (int i = 0; i < k; ++i) {
          if (i + j + k < N && (i + j) / (2 * p) == (i + j + k) / (2 * p)) {
            const float a = e[i + j], b = e[i + j + k];
            e[i + j] = a < b ? a : b;
            e[i + j + k] = a < b ? b : a;
          }
        }
      }
    }
  }
  if (rules == BLCD_RASTER_PIL9) {
#pragma unroll
    for (int k = 1; k < N; k += 2)
      if (k < cnt) mask |= span_mask(round_up_px_i(e[k - 1]), round_down_px_i(e[k]), w, x_off);
    return mask;
  }
  bool have_pos = false;
  int pos = 0;
#pragma unroll
  for (int k = 1; k < N; k += 2) {
    if (k < cnt) {
      int xs = round_up_px_i(e[k - 1]);
      const int xe = round_down_px_i(e[k]);
      const bool skip = have_pos && xe < pos;
      if (have_pos && xs < pos) xs = pos;
      if (!skip && xe >= xs) {
        mask |= span_mask(xs, xe, w, x_off);
        pos = xe + 1; have_pos = true;
      }
    }
  }
  // apex extension (same tests as polygon_row; slopes from the table)
#pragma unroll
  for (int i = 1; i < NE; ++i) {
    if (i >= ne) continue;
    const bool ilast = (i + 1 >= n);
    const int cx0 = P.x[i], cy0 = P.y[i];
    const int cx1 = (ilast || i + 1 >= NE) ? P.x[0] : P.x[i + 1 < NE ? i + 1 : 0], cy1 = (ilast || i + 1 >= NE) ? P.y[0] : P.y[i + 1 < NE ? i + 1 : 0];
    if (cy0 == cy1 || cx0 == cx1) continue;
    const int cmin = cy0 < cy1 ? cy0 : cy1, cmax = cy0 < cy1 ? cy1 : cy0;
    if (y != cmin && y != cmax) continue;
    const int vx = (cy0 == y) ? cx0 : cx1;
    const bool cpos = ((cx1 - cx0) > 0) == ((cy1 - cy0) > 0);
#pragma unroll
    for (int k = 0; k < i; ++k) {
      const int ox0 = P.x[k], oy0 = P.y[k], ox1 = P.x[k + 1], oy1 = P.y[k + 1];   // k + 1 <= i < n: never the closing edge
      if (oy0 == oy1 || ox0 == ox1) continue;
      const bool opos = ((ox1 - ox0) > 0) == ((oy1 - oy0) > 0);
      if (cpos != opos) continue;
      const int omin = oy0 < oy1 ? oy0 : oy1, omax = oy0 < oy1 ? oy1 : oy0;
      const bool top = (y == cmin && y == omin);
      const bool bot = (y == cmax && y == omax && y == Ymax);
      if (!top && !bot) continue;
      const int ovx = (oy0 == y) ? ox0 : ox1;
      if (vx != ovx) continue;
      const int y2 = (y == Ymax) ? y - 1 : y + 1;
      const float a1 = edge_x_at(cx0, cy0, S.dx[i], y2), a2 = edge_x_at(ox0, oy0, S.dx[k], y2);
      if ((bot && cpos) || (top && !cpos)) {
        const int s = round_up_px_i(BLCD_FADD(a1 > a2 ? a1 : a2, 1.0f));
        if (s <= vx) mask |= span_mask(s, vx, w, x_off);
      } else {
        const int t = round_up_px_i(a1 < a2 ? a1 : a2) - 1;
        if (t >= vx) mask |= span_mask(vx, t, w, x_off);
      }
    }
  }
  return mask;
}

BLCD_HD RowMask polygon_row_fast(const PolyPx& P, const PolySlopes& S, int y, int w, int h, int rules, int ylo, int yhi, int x_off = 0) {
  // boxes and the luxo head (every limb of every robot): the unrolled 4-edge form.  Shapes with more vertices (one hull per crab /
  // walker) take the general row function: unrolling 8 edges would double the kernel's registers for one body per frame.
  return P.n <= 4 ? polygon_row_t<4>(P, S, y, w, h, rules, ylo, yhi, x_off) : polygon_row(P, y, w, h, rules, ylo, yhi, x_off);
}

// integer pixel vertices of a polygon fixture under transform (px, py, s, c)
BLCD_HD void polygon_px(PolyPx& out, const DShape& sh, float px, float py, float s, float c, double world_w, double lcd_w) {
  out.n = sh.count;
  for (int i = 0; i < BLCD_MAX_VERTS; ++i) {
    if (i < sh.count) {
      float vx = sh.v[i].x, vy = sh.v[i].y;
      float wx = BLCD_FADD(BLCD_FSUB(BLCD_FMUL(c, vx), BLCD_FMUL(s, vy)), px);
      float wy = BLCD_FADD(BLCD_FADD(BLCD_FMUL(s, vx), BLCD_FMUL(c, vy)), py);
      out.x[i] = to_px((double)wx, world_w, lcd_w);
      out.y[i] = to_px((double)wy, world_w, lcd_w);
    }
  }
}

// ink mask of canvas row y (y-up canvas coordinate) for one body
// lcd_w scales metres to pixels; the ink returned is that of columns [x_off, x_off + win_w) (win_w = 0: the whole frame)
BLCD_HD RowMask body_row(const DShape& sh, float px, float py, float s, float c, int y, int world_w, int lcd_w, int lcd_h, int rules,
                         int x_off = 0, int win_w = 0) {
  const int cw = win_w > 0 ? win_w : lcd_w;
  const double ww = (double)world_w, lw = (double)lcd_w;
  if (sh.type == SH_CIRCLE) {
    const double r = (double)sh.radius;
    return ellipse_row(to_px((double)px - r, ww, lw), to_px((double)py - r, ww, lw), to_px((double)px + r, ww, lw),
                       to_px((double)py + r, ww, lw), y, cw, x_off);
  }
  PolyPx P;
  polygon_px(P, sh, px, py, s, c, ww, lw);
  return polygon_row(P, y, cw, lcd_h, rules, 1, 0, x_off);
}

// Per-body raster setup, computed once per frame: integer box of a circle, or integer vertices of a polygon.  The
// per-row functions above then only do integer / fp32 scanline work.
struct BodyPx {
  int kind;          // SH_CIRCLE or SH_POLY
  int x0, y0, x1, y1;
  PolyPx P;
};

BLCD_HD void body_px(BodyPx& o, const DShape& sh, float px, float py, float s, float c, int world_w, int lcd_w) {
  const double ww = (double)world_w, lw = (double)lcd_w;
  o.kind = sh.type;
  if (sh.type == SH_CIRCLE) {
    const double r = (double)sh.radius;
    o.x0 = to_px((double)px - r, ww, lw); o.y0 = to_px((double)py - r, ww, lw);
    o.x1 = to_px((double)px + r, ww, lw); o.y1 = to_px((double)py + r, ww, lw);
    o.P.n = 0;
  } else {
    polygon_px(o.P, sh, px, py, s, c, ww, lw);
    int ylo = o.P.y[0], yhi = o.P.y[0];
    for (int i = 1; i < BLCD_MAX_VERTS; ++i)
      if (i < o.P.n) { ylo = o.P.y[i] < ylo ? o.P.y[i] : ylo; yhi = o.P.y[i] > yhi ? o.P.y[i] : yhi; }
    o.y0 = ylo; o.y1 = yhi; o.x0 = 0; o.x1 = 0;
  }
}

// the converged-warp variant (k_render_bodies): slopes once per body, polygon_row_t
struct BodyPxFast {
  BodyPx b;
  PolySlopes S;
};
BLCD_HD void body_px_fast(BodyPxFast& o, const DShape& sh, float px, float py, float s, float c, int world_w, int lcd_w) {
  body_px(o.b, sh, px, py, s, c, world_w, lcd_w);
  if (o.b.kind != SH_CIRCLE) polygon_slopes(o.S, o.b.P);
}
BLCD_HD RowMask body_px_row_fast(const BodyPxFast& o, int y, int win_w, int lcd_h, int rules, int x_off = 0) {
  if (y < o.b.y0 || y > o.b.y1) return 0u;
  if (o.b.kind == SH_CIRCLE) return ellipse_row(o.b.x0, o.b.y0, o.b.x1, o.b.y1, y, win_w, x_off);
  return polygon_row_fast(o.b.P, o.S, y, win_w, lcd_h, rules, o.b.y0, o.b.y1, x_off);
}

// win_w / x_off: the column window whose ink is returned (the whole frame when it fits one RowMask)
BLCD_HD RowMask body_px_row(const BodyPx& o, int y, int win_w, int lcd_h, int rules, int x_off = 0) {
  if (y < o.y0 || y > o.y1) return 0u;   // outside the shape's rows: neither the ellipse nor the polygon rules draw anything
  if (o.kind == SH_CIRCLE) return ellipse_row(o.x0, o.y0, o.x1, o.y1, y, win_w, x_off);
  return polygon_row(o.P, y, win_w, lcd_h, rules, o.y0, o.y1, x_off);
}

BLCD_HD RowMask row_bits_from_ink(RowMask ink, int lcd_w) {
  RowMask full = lcd_w >= kRowBits ? ~(RowMask)0 : (((RowMask)1 << lcd_w) - 1u);
  return (~ink) & full;
}

// output words per frame row (include/boxlcd_b200.h: word k = pixels 32k .. 32k+31) and word k of a row
BLCD_HD int row_words(int lcd_w) { return BLCD_LCD_WORDS(lcd_w); }
BLCD_HD uint32_t row_word(RowMask bits, int k) { return sizeof(RowMask) == 4 ? (uint32_t)bits : (uint32_t)((uint64_t)bits >> (32 * k)); }

}  // namespace BLCD_NS
