/*
 * lcd_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement of the reference's LCD renderer, `WorldEnv.lcd_render` (boxLCD/world_env.py:460-512), i.e. of
 * what Pillow's C rasterizer (`ImageDraw.ellipse`, `ImageDraw.polygon`, `Image.transpose(FLIP_TOP_BOTTOM)`) does
 * with the coordinates the reference hands it.  Pillow is a third-party dependency of the reference
 * (requirements.txt:19 pins Pillow==9.0.1; this image has 12.2.0) and its sources are not under /root/reference, so the
 * fill rules are restated from its published algorithm (libImaging/Draw.c: `polygon_generic`, `ellipseNew`) and PINNED
 * against the unmodified reference `lcd_render` executed on this container's Pillow 12.2.0
 * (tests/golden/make_lcd_golden.py -> tests/golden/lcd_*.npz; tests/test_lcd_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may call into this file.
 *
 * Rule sets: 0 = "pil12" (pinned here against this image's Pillow), 1 = "pil9" (the Pillow of the reference's recorded
 * episodes: no apex extension, no span-overlap bookkeeping, horizontal edges not drawn; checked against the recorded robot
 * gifs, tests/test_gif_episodes.py -- 95 % of their frames bit-exact, not pinned pixel for pixel).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LCD_MAX_VERTS 8
#define LCD_MAX_W 256

typedef struct {
  int32_t kind;  /* 0 circle, else polygon */
  int32_t n;     /* polygon vertex count (stored hull order) */
  float radius;
  float verts[LCD_MAX_VERTS][2];
} lcd_shape;

typedef struct {
  int w, h;
  uint8_t ink[LCD_MAX_W * LCD_MAX_W];
} canvas;

/* Draw.c ROUND_UP / ROUND_DOWN (round half away from / toward the span) */
static int round_up(float f) { return f >= 0.0f ? (int)floor((double)f + 0.5) : -(int)floor(-(double)f + 0.5); }
static int round_down(float f) { return f >= 0.0f ? (int)ceil((double)f - 0.5) : -(int)ceil(-(double)f - 0.5); }

/* Draw.c hline8: inclusive span, ends swapped if inverted, clipped to the image */
static void hline(canvas* cv, int x0, int y, int x1) {
  if (y < 0 || y >= cv->h) return;
  if (x0 > x1) { int t = x0; x0 = x1; x1 = t; }
  if (x0 < 0) x0 = 0;
  if (x1 >= cv->w) x1 = cv->w - 1;
  for (int x = x0; x <= x1; ++x) cv->ink[y * cv->w + x] = 1;
}

/* Draw.c ellipseNew / quarter_init + quarter_next, filled, width 0: walk the first-quadrant boundary of the ellipse
 * inscribed in the integer box in doubled coordinates, keep the largest X per Y, mirror to the four quadrants. */
static int64_t ell_delta(int64_t a, int64_t b, int64_t x, int64_t y) {
  int64_t d = a * a * y * y + b * b * x * x - a * a * b * b;
  return d < 0 ? -d : d;
}

static void fill_ellipse(canvas* cv, int x0, int y0, int x1, int y1) {
  int a = x1 - x0, b = y1 - y0;
  if (a < 0 || b < 0 || (a == 0 && b == 0)) return;
  int* xmax = (int*)malloc(sizeof(int) * (size_t)(b + 3));
  for (int i = 0; i < b + 3; ++i) xmax[i] = -1;
  int cx = a, cy = b % 2;
  const int ex = a % 2, ey = b;
  for (;;) {
    if (cx > xmax[cy]) xmax[cy] = cx;
    if (cx == ex && cy == ey) break;
    int nx = cx, ny = cy + 2;
    int64_t nd = ell_delta(a, b, nx, ny);
    if (cx > 1) {
      int64_t d1 = ell_delta(a, b, cx - 2, cy + 2);
      if (nd > d1) { nx = cx - 2; ny = cy + 2; nd = d1; }
      int64_t d2 = ell_delta(a, b, cx - 2, cy);
      if (nd > d2) { nx = cx - 2; ny = cy; }
    }
    cx = nx; cy = ny;
  }
  for (int Y = b % 2; Y <= b; Y += 2) {
    if (xmax[Y] < 0) continue;
    int X = xmax[Y];
    hline(cv, x0 + (a - X) / 2, y0 + (Y + b) / 2, x0 + (a + X) / 2);
    hline(cv, x0 + (a - X) / 2, y0 + (-Y + b) / 2, x0 + (a + X) / 2);
  }
  free(xmax);
}

typedef struct {
  int x0, y0, x1, y1, ymin, ymax;
  float dx;
} edge;

static float edge_x(const edge* e, int y) {
  /* fp32, separately rounded multiply and add (Draw.c: (ymin - y0) * dx + x0 with float dx; built without FMA) */
  volatile float m = (float)(y - e->y0) * e->dx;
  volatile float r = m + (float)e->x0;
  return r;
}

static int cmp_float(const void* a, const void* b) {
  float x = *(const float*)a, y = *(const float*)b;
  return (x > y) - (x < y);
}

/* Draw.c polygon_generic on integer vertices */
static void fill_polygon(canvas* cv, const int (*P)[2], int n, int rules) {
  edge all[LCD_MAX_VERTS + 1], tab[LCD_MAX_VERTS + 1];
  int n_all = 0, n_tab = 0;
  for (int i = 0; i + 1 < n; ++i) {
    edge e = {P[i][0], P[i][1], P[i + 1][0], P[i + 1][1], 0, 0, 0.0f};
    all[n_all++] = e;
  }
  if (n > 0 && (P[n - 1][0] != P[0][0] || P[n - 1][1] != P[0][1])) {
    edge e = {P[n - 1][0], P[n - 1][1], P[0][0], P[0][1], 0, 0, 0.0f};
    all[n_all++] = e;
  }
  if (n_all == 0) return;
  int ylo = all[0].y0, yhi = all[0].y0;
  for (int i = 0; i < n_all; ++i) {
    edge* e = &all[i];
    e->ymin = e->y0 < e->y1 ? e->y0 : e->y1;
    e->ymax = e->y0 < e->y1 ? e->y1 : e->y0;
    if (e->ymin < ylo) ylo = e->ymin;
    if (e->ymax > yhi) yhi = e->ymax;
    if (e->y0 == e->y1) {
      /* current Pillow draws horizontal edges as spans of their own; the Pillow behind the reference's recorded episodes
       * (requirements.txt pins 9.0.1) did not: a limb lying flat inside one pixel row is invisible there */
      if (rules != 1) hline(cv, e->x0, e->y0, e->x1);
      continue;
    }
    e->dx = (float)(e->x1 - e->x0) / (float)(e->y1 - e->y0);
    tab[n_tab++] = *e;
  }
  const int Ymin = ylo > 0 ? ylo : 0, Ymax = yhi < cv->h ? yhi : cv->h;
  float xx[2 * (LCD_MAX_VERTS + 1)];
  for (int y = Ymin; y <= Ymax; ++y) {
    int j = 0;
    for (int i = 0; i < n_tab; ++i) {
      const edge* e = &tab[i];
      if (e->ymin <= y && y <= e->ymax) {
        xx[j++] = edge_x(e, y);
        if (y == e->ymax && y < Ymax) { xx[j] = xx[j - 1]; ++j; }
      }
    }
    qsort(xx, (size_t)j, sizeof(float), cmp_float);
    if (rules == 1) { /* pil9: every pair drawn, inverted spans swapped by hline */
      for (int k = 1; k < j; k += 2) hline(cv, round_up(xx[k - 1]), y, round_down(xx[k]));
      continue;
    }
    int have_pos = 0, pos = 0;
    for (int k = 1; k < j; k += 2) {
      int xs = round_up(xx[k - 1]), xe = round_down(xx[k]);
      if (have_pos) {
        if (xe < pos) continue;
        if (xs < pos) xs = pos;
      }
      if (xe < xs) continue;
      hline(cv, xs, y, xe);
      pos = xe + 1; have_pos = 1;
    }
    /* apex extension: two same-direction sloped edges meeting at an integer vertex on this row */
    for (int i = 0; i < n_tab; ++i) {
      const edge* c = &tab[i];
      for (int k = 0; k < i; ++k) {
        const edge* o = &tab[k];
        if (c->dx == 0.0f || o->dx == 0.0f || (c->dx > 0.0f) != (o->dx > 0.0f)) continue;
        int top = (y == c->ymin && y == o->ymin);
        int bot = (y == c->ymax && y == o->ymax && y == Ymax);
        if (!top && !bot) continue;
        int vx = (c->y0 == y) ? c->x0 : c->x1;
        int ox = (o->y0 == y) ? o->x0 : o->x1;
        if (vx != ox) continue;
        int y2 = (y == Ymax) ? y - 1 : y + 1;
        float a1 = edge_x(c, y2), a2 = edge_x(o, y2);
        if ((bot && c->dx > 0.0f) || (top && c->dx < 0.0f)) {
          volatile float m = (a1 > a2 ? a1 : a2) + 1.0f;
          int s = round_up(m);
          if (s <= vx) hline(cv, s, y, vx);
        } else {
          int t = round_up(a1 < a2 ? a1 : a2) - 1;
          if (t >= vx) hline(cv, vx, y, t);
        }
      }
    }
  }
}

/* world metres -> canvas pixels: fp64 divide by WIDTH, multiply by the frame width, truncate toward zero
 * (world_env.py:493-505; PIL copies the doubles into an int array) */
static int to_px(double v, int world_w, int lcd_w) { return (int)(v / (double)world_w * (double)lcd_w); }

/* poses: [n, n_bodies, 4] = (px, py, sin, cos) fp32, the b2Transform of every dynamic body in draw order.
 * shapes: [n_bodies] (or [n, n_bodies] when per_world_shapes != 0).  bits: [n, lcd_h] uint32 ([n, lcd_h, 2] for 32 < lcd_w <= 64), bit x = pixel x,
 * 1 = background, row 0 = top of the world. */
void blcd_oracle_lcd(const lcd_shape* shapes, int per_world_shapes, int n_bodies, const float* poses, int64_t n, int world_w,
                     int lcd_w, int lcd_h, int rules, uint32_t* bits) {
  canvas cv;
  cv.w = lcd_w; cv.h = lcd_h;
  for (int64_t w = 0; w < n; ++w) {
    memset(cv.ink, 0, (size_t)(lcd_w * lcd_h));
    for (int b = 0; b < n_bodies; ++b) {
      const lcd_shape* sh = &shapes[(per_world_shapes ? w * n_bodies : 0) + b];
      const float* p = &poses[(w * n_bodies + b) * 4];
      const float px = p[0], py = p[1], s = p[2], c = p[3];
      if (sh->kind == 0) {
        const double r = (double)sh->radius;
        fill_ellipse(&cv, to_px((double)px - r, world_w, lcd_w), to_px((double)py - r, world_w, lcd_w),
                     to_px((double)px + r, world_w, lcd_w), to_px((double)py + r, world_w, lcd_w));
      } else {
        int P[LCD_MAX_VERTS][2];
        for (int i = 0; i < sh->n; ++i) {
          /* b2Mul(b2Transform, b2Vec2): fp32 with every product and sum rounded separately (no FMA) */
          volatile float cx = c * sh->verts[i][0], sy = s * sh->verts[i][1];
          volatile float sx = s * sh->verts[i][0], cy = c * sh->verts[i][1];
          volatile float rx = cx - sy, ry = sx + cy;
          volatile float wx = rx + px, wy = ry + py;
          P[i][0] = to_px((double)wx, world_w, lcd_w);
          P[i][1] = to_px((double)wy, world_w, lcd_w);
        }
        fill_polygon(&cv, (const int (*)[2])P, sh->n, rules);
      }
    }
    for (int R = 0; R < lcd_h; ++R) {
      const uint8_t* row = &cv.ink[(lcd_h - 1 - R) * lcd_w];
      const int lw = (lcd_w + 31) / 32;   /* words per row: word k = pixels 32k .. 32k+31 */
      for (int k = 0; k < lw; ++k) {
        uint32_t m = 0;
        for (int x = 32 * k; x < lcd_w && x < 32 * k + 32; ++x) m |= (uint32_t)(row[x] == 0) << (x - 32 * k);
        bits[((size_t)w * lcd_h + R) * lw + k] = m;
      }
    }
  }
}
