"""Evaluation plots without matplotlib (eval.py:270-469 of the reference: position tracking and tracking error
of the best / worst episode, written as PNG files under ``<output_dir>/plots``).

matplotlib is not part of this image, and the plots are host-side report output, not part of the accelerated path:
this module is a small retained-mode line-chart renderer on a numpy canvas (lines, dashed lines, alpha bands,
scatter marks, ticks, grid, legend, 5x7 bitmap text) with a zlib PNG encoder.  ``Figure`` / ``Axes`` carry the
handful of methods the reference's plotting code uses (``plot``, ``fill_between``, ``scatter``, ``axhline``,
``set_xlabel`` / ``set_ylabel`` / ``set_title``, ``legend``, ``grid``, ``set_ylim``, ``suptitle``, ``savefig``), so
``Evaluator.plot_trajectory`` / ``plot_tracking_error`` / ``generate_all_plots`` keep the reference's signatures,
file names and figure sizes (12 x 10 and 12 x 5 inches at 150 dpi).
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

# the classic public-domain 5x7 LCD font, ASCII 32..126: five column bytes per glyph, bit 0 = top row
_FONT = bytes.fromhex(
    "0000000000" "00005f0000" "0007000700" "147f147f14" "242a7f2a12" "2313086462" "3649552250" "0005030000"
    "001c224100" "0041221c00" "14083e0814" "08083e0808" "0050300000" "0808080808" "0060600000" "2010080402"
    "3e5149453e" "00427f4000" "4261514946" "2141454b31" "1814127f10" "2745454539" "3c4a494930" "0171090503"
    "3649494936" "064949291e" "0036360000" "0056360000" "0814224100" "1414141414" "0041221408" "0201510906"
    "324979413e" "7e1111117e" "7f49494936" "3e41414122" "7f4141221c" "7f49494941" "7f09090901" "3e4149497a"
    "7f0808087f" "00417f4100" "2040413f01" "7f08142241" "7f40404040" "7f020c027f" "7f0408107f" "3e4141413e"
    "7f09090906" "3e4151215e" "7f09192946" "4649494931" "01017f0101" "3f4040403f" "1f2040201f" "3f4038403f"
    "6314081463" "0708700807" "6151494543" "007f414100" "0204081020" "0041417f00" "0402010204" "4040404040"
    "0001020400" "2054545478" "7f48444438" "3844444420" "384444487f" "3854545418" "087e090102" "0c5252523e"
    "7f08040478" "00447d4000" "2040443d00" "7f10284400" "00417f4000" "7c04180478" "7c08040478" "3844444438"
    "7c14141408" "081414187c" "7c08040408" "4854545420" "043f444020" "3c4040207c" "1c2040201c" "3c4030403c"
    "4428102844" "0c5050503c" "4464544c44" "0008364100" "00007f0000" "0041360800" "1008081008")

COLORS = {
    "tab:blue": (31, 119, 180), "tab:orange": (255, 127, 14), "tab:green": (44, 160, 44), "red": (255, 0, 0),
    "r": (255, 0, 0), "green": (0, 128, 0), "g": (0, 128, 0), "blue": (0, 0, 255), "b": (0, 0, 255),
    "black": (0, 0, 0), "k": (0, 0, 0), "gray": (128, 128, 128),
}


def _rgb(color) -> np.ndarray:
    if isinstance(color, str):
        return np.array(COLORS[color], dtype=np.float32)
    return np.asarray(color, dtype=np.float32)[:3]


def text_bitmap(text: str, scale: int = 2) -> np.ndarray:
    """Boolean [7 * scale, 6 * scale * len(text)] bitmap of ``text`` (characters outside ASCII 32..126 print '?')."""
    cols = []
    for ch in text:
        k = ord(ch) - 32
        if not 0 <= k < 95:
            k = ord("?") - 32
        cols.extend(_FONT[5 * k:5 * k + 5])
        cols.append(0)
    glyph = ((np.array(cols, dtype=np.uint8)[None, :] >> np.arange(7, dtype=np.uint8)[:, None]) & 1).astype(bool)
    return np.kron(glyph, np.ones((scale, scale), dtype=bool))


def nice_ticks(lo: float, hi: float, target: int = 7) -> np.ndarray:
    """Tick positions at 1 / 2 / 5 x 10^k spacing inside [lo, hi]."""
    if not np.isfinite(lo) or not np.isfinite(hi) or hi <= lo:
        return np.array([lo])
    raw = (hi - lo) / max(target, 1)
    mag = 10.0 ** np.floor(np.log10(raw))
    step = next(m * mag for m in (1.0, 2.0, 5.0, 10.0) if m * mag >= raw)
    first = np.ceil(lo / step - 1e-9)
    last = np.floor(hi / step + 1e-9)
    return np.arange(first, last + 1) * step


def _tick_label(v: float, step: float) -> str:
    if abs(v) < 1e-9 * step:
        v = 0.0                      # no "-0"
    if step >= 1.0:
        return f"{v:.0f}"
    return f"{v:.{min(6, int(np.ceil(-np.log10(step) - 1e-9)))}f}"


class Canvas:
    """RGB float canvas, origin at the top left."""

    def __init__(self, width: int, height: int):
        self.w, self.h = int(width), int(height)
        self.px = np.full((self.h, self.w, 3), 255.0, dtype=np.float32)

    def blend(self, ys, xs, color, alpha: float = 1.0) -> None:
        ys = np.asarray(ys).astype(np.int64).ravel()
        xs = np.asarray(xs).astype(np.int64).ravel()
        ok = (ys >= 0) & (ys < self.h) & (xs >= 0) & (xs < self.w)
        ys, xs = ys[ok], xs[ok]
        c = _rgb(color)
        if alpha >= 1.0:
            self.px[ys, xs] = c
        else:
            flat = np.unique(ys * self.w + xs)   # each pixel once, however many marks overlap there
            ys, xs = flat // self.w, flat % self.w
            self.px[ys, xs] = (1.0 - alpha) * self.px[ys, xs] + alpha * c

    def rect(self, x0, y0, x1, y1, color, alpha: float = 1.0) -> None:
        x0, x1 = sorted((int(round(x0)), int(round(x1))))
        y0, y1 = sorted((int(round(y0)), int(round(y1))))
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, self.w), min(y1, self.h)
        if x1 > x0 and y1 > y0:
            self.px[y0:y1, x0:x1] = (1.0 - alpha) * self.px[y0:y1, x0:x1] + alpha * _rgb(color)

    def polyline(self, xs, ys, color, width: int = 2, dash: tuple[int, int] | None = None, clip=None,
                 alpha: float = 1.0) -> None:
        """Connected segments through pixel coordinates (xs, ys); ``dash`` = (on, off) lengths in pixels;
        ``clip`` = (x0, y0, x1, y1) pixel box."""
        xs = np.asarray(xs, dtype=np.float64)
        ys = np.asarray(ys, dtype=np.float64)
        good = np.isfinite(xs) & np.isfinite(ys)
        if good.sum() < 2:
            return
        seg = good[:-1] & good[1:]
        x0, y0, x1, y1 = xs[:-1][seg], ys[:-1][seg], xs[1:][seg], ys[1:][seg]
        n = np.maximum(np.ceil(np.maximum(np.abs(x1 - x0), np.abs(y1 - y0))).astype(np.int64), 1)
        n = np.minimum(n, 4 * (self.w + self.h))       # a segment far outside the canvas needs no finer sampling
        idx = np.repeat(np.arange(len(n)), n)
        start = np.cumsum(n) - n
        t = (np.arange(n.sum()) - start[idx]) / n[idx]
        px = x0[idx] + t * (x1 - x0)[idx]
        py = y0[idx] + t * (y1 - y0)[idx]
        px, py = np.append(px, x1[-1]), np.append(py, y1[-1])
        if dash is not None:
            d = np.hypot(np.diff(px, prepend=px[0]), np.diff(py, prepend=py[0])).cumsum()
            on = (d % (dash[0] + dash[1])) < dash[0]
            px, py = px[on], py[on]
        if clip is not None:
            inside = (px >= clip[0]) & (px <= clip[2]) & (py >= clip[1]) & (py <= clip[3])
            px, py = px[inside], py[inside]
        ix, iy = np.round(px).astype(np.int64), np.round(py).astype(np.int64)
        offs = range(-((width - 1) // 2), width // 2 + 1)
        for dy in offs:
            for dx in offs:
                self.blend(iy + dy, ix + dx, color, alpha)

    def marks(self, xs, ys, color, size: int = 3, alpha: float = 1.0, clip=None) -> None:
        xs, ys = np.asarray(xs, dtype=np.float64), np.asarray(ys, dtype=np.float64)
        good = np.isfinite(xs) & np.isfinite(ys)
        if clip is not None:
            good &= (xs >= clip[0]) & (xs <= clip[2]) & (ys >= clip[1]) & (ys <= clip[3])
        ix, iy = np.round(xs[good]).astype(np.int64), np.round(ys[good]).astype(np.int64)
        offs = np.arange(-(size // 2), size - size // 2)
        dy, dx = np.meshgrid(offs, offs, indexing="ij")
        self.blend((iy[:, None] + dy.ravel()[None, :]), (ix[:, None] + dx.ravel()[None, :]), color, alpha)

    def text(self, x: float, y: float, s: str, scale: int = 2, color="black", anchor: str = "lt",
             vertical: bool = False) -> None:
        """``anchor``: horizontal l / c / r then vertical t / c / b of the text box relative to (x, y)."""
        bm = text_bitmap(s, scale)
        if vertical:
            bm = np.rot90(bm)
        h, w = bm.shape
        x0 = x - {"l": 0, "c": w / 2, "r": w}[anchor[0]]
        y0 = y - {"t": 0, "c": h / 2, "b": h}[anchor[1]]
        yy, xx = np.nonzero(bm)
        self.blend(yy + int(round(y0)), xx + int(round(x0)), color)

    def png_bytes(self) -> bytes:
        img = np.clip(np.round(self.px), 0, 255).astype(np.uint8)
        raw = np.concatenate([np.zeros((self.h, 1), dtype=np.uint8), img.reshape(self.h, self.w * 3)], axis=1).tobytes()

        def chunk(tag: bytes, data: bytes) -> bytes:
            return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

        return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", self.w, self.h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def read_png(path) -> np.ndarray:
    """Decode a PNG written by ``Canvas.png_bytes`` (8-bit RGB, filter 0) back into [H, W, 3] uint8 — for tests."""
    data = Path(path).read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n", "not a PNG file"
    pos, idat, w, h = 8, b"", 0, 0
    while pos < len(data):
        n, tag = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + body) & 0xffffffff
        if tag == b"IHDR":
            w, h, depth, ctype = struct.unpack(">IIBB", body[:10])
            assert (depth, ctype) == (8, 2)
        elif tag == b"IDAT":
            idat += body
        pos += 12 + n
    rows = np.frombuffer(zlib.decompress(idat), dtype=np.uint8).reshape(h, 1 + 3 * w)
    assert not rows[:, 0].any()
    return rows[:, 1:].reshape(h, w, 3)


class Axes:
    """One chart: the artists are recorded and drawn by ``Figure.savefig``."""

    def __init__(self):
        self.artists: list[tuple] = []
        self.xlabel = self.ylabel = self.title = ""
        self.show_legend = self.show_grid = False
        self.ylim: list[float | None] = [None, None]

    # -- the calls eval.py makes -----------------------------------------------------------------
    def plot(self, x, y, fmt: str | None = None, label: str = "", color=None, linestyle: str = "-",
             linewidth: float = 1.0, **_):
        if fmt:                      # "b-", "r--", ...
            color = color or fmt[0]
            linestyle = fmt[1:] or linestyle
        self.artists.append(("line", np.asarray(x, float), np.asarray(y, float), color or "tab:blue", linestyle, label,
                             linewidth))

    def fill_between(self, x, y1, y2, alpha: float = 0.2, color="green", label: str = "", **_):
        self.artists.append(("band", np.asarray(x, float), np.asarray(y1, float), np.asarray(y2, float), color, alpha,
                             label))

    def scatter(self, x, y, c="green", s: float = 2, alpha: float = 0.5, label: str = "", **_):
        self.artists.append(("marks", np.asarray(x, float), np.asarray(y, float), c, alpha, label, s))

    def axhline(self, y: float, color="r", linestyle: str = "--", label: str = "", **_):
        self.artists.append(("hline", float(y), color, linestyle, label))

    def set_xlabel(self, s: str):
        self.xlabel = s

    def set_ylabel(self, s: str):
        self.ylabel = s

    def set_title(self, s: str):
        self.title = s

    def get_xlabel(self) -> str:
        return self.xlabel

    def get_ylabel(self) -> str:
        return self.ylabel

    def get_title(self) -> str:
        return self.title

    def legend(self, loc: str = "upper right", **_):
        self.show_legend = True

    def grid(self, on: bool = True, **_):
        self.show_grid = bool(on)

    def set_ylim(self, bottom=None, top=None):
        if bottom is not None:
            self.ylim[0] = float(bottom)
        if top is not None:
            self.ylim[1] = float(top)

    # -- drawing ----------------------------------------------------------------------------------
    def data_limits(self):
        xs, ys = [], []
        for a in self.artists:
            if a[0] in ("line", "marks"):
                xs.append(a[1]), ys.append(a[2])
            elif a[0] == "band":
                xs.append(a[1]), ys.append(a[2]), ys.append(a[3])
            elif a[0] == "hline":
                ys.append(np.array([a[1]]))

        def span(parts):
            v = np.concatenate([np.ravel(p) for p in parts]) if parts else np.array([0.0, 1.0])
            v = v[np.isfinite(v)]
            if v.size == 0:
                return 0.0, 1.0
            lo, hi = float(v.min()), float(v.max())
            if hi <= lo:
                pad = 0.5 if lo == 0.0 else 0.05 * abs(lo)
                return lo - pad, hi + pad
            return lo, hi

        x0, x1 = span(xs)
        y0, y1 = span(ys)
        pad = 0.05 * (y1 - y0)
        y0, y1 = y0 - pad, y1 + pad
        if self.ylim[0] is not None:
            y0 = self.ylim[0]
        if self.ylim[1] is not None:
            y1 = self.ylim[1]
        if y1 <= y0:
            y1 = y0 + 1.0
        return x0, x1, y0, y1

    def draw(self, cv: Canvas, box, x_range=None, font: int = 2, x_tick_labels: bool = True) -> None:
        L, T, R, B = box
        x0, x1, y0, y1 = self.data_limits()
        if x_range is not None:
            x0, x1 = x_range
        sx, sy = (R - L) / (x1 - x0), (B - T) / (y1 - y0)

        def X(v):
            return L + (np.asarray(v, float) - x0) * sx

        def Y(v):
            return B - (np.asarray(v, float) - y0) * sy

        clip = (L, T, R, B)
        xt, yt = nice_ticks(x0, x1, 8), nice_ticks(y0, y1, 6)
        if self.show_grid:
            for v in xt:
                cv.polyline([X(v), X(v)], [T, B], (215, 215, 215), 1)
            for v in yt:
                cv.polyline([L, R], [Y(v), Y(v)], (215, 215, 215), 1)
        legend = []
        # bands below lines below marks (matplotlib draws patches / collections under lines), legend in call order
        order = {"band": 0, "line": 1, "hline": 1, "marks": 2}
        legend_slots = {id(a): i for i, a in enumerate(self.artists)}
        for a in sorted(self.artists, key=lambda a: order[a[0]]):
            if a[0] == "band":
                _, x, lo, hi, color, alpha, label = a
                cols = np.arange(int(np.ceil(max(X(x.min()), L))), int(np.floor(min(X(x.max()), R))) + 1)
                if cols.size:
                    xd = x0 + (cols - L) / sx
                    top = np.clip(np.round(Y(np.interp(xd, x, hi))), T, B).astype(np.int64)
                    bot = np.clip(np.round(Y(np.interp(xd, x, lo))), T, B).astype(np.int64)
                    n = np.maximum(bot - top + 1, 0)
                    idx = np.repeat(np.arange(cols.size), n)
                    rows = top[idx] + (np.arange(n.sum()) - (np.cumsum(n) - n)[idx])
                    cv.blend(rows, cols[idx], color, alpha)
                if label:
                    legend.append((legend_slots[id(a)], "band", color, alpha, label))
            elif a[0] == "line":
                _, x, y, color, style, label, lw = a
                dash = (12, 8) if style == "--" else None
                cv.polyline(X(x), Y(y), color, 2 if lw <= 1.5 else 3, dash, clip)
                if label:
                    legend.append((legend_slots[id(a)], "line", color, dash, label))
            elif a[0] == "marks":
                _, x, y, color, alpha, label, s = a
                cv.marks(X(x), Y(y), color, 5 if s <= 4 else 7, alpha, clip)
                if label:
                    legend.append((legend_slots[id(a)], "marks", color, alpha, label))
            elif a[0] == "hline":
                _, v, color, style, label = a
                dash = (12, 8) if style == "--" else None
                if y0 <= v <= y1:
                    cv.polyline([L, R], [Y(v), Y(v)], color, 2, dash, clip)
                if label:
                    legend.append((legend_slots[id(a)], "line", color, dash, label))
        # frame, ticks, labels
        cv.polyline([L, R, R, L, L], [T, T, B, B, T], "black", 2)
        xstep = xt[1] - xt[0] if len(xt) > 1 else 1.0
        ystep = yt[1] - yt[0] if len(yt) > 1 else 1.0
        for v in xt:
            cv.polyline([X(v), X(v)], [B, B + 8], "black", 2)
            if x_tick_labels:
                cv.text(float(X(v)), B + 12, _tick_label(v, xstep), font, anchor="ct")
        for v in yt:
            cv.polyline([L - 8, L], [Y(v), Y(v)], "black", 2)
            cv.text(L - 12, float(Y(v)), _tick_label(v, ystep), font, anchor="rc")
        if self.xlabel:
            cv.text((L + R) / 2, B + 16 + 7 * font, self.xlabel, font, anchor="ct")
        if self.ylabel:
            cv.text(L - 40 - 6 * font * 6, (T + B) / 2, self.ylabel, font, anchor="cc", vertical=True)
        if self.title:
            cv.text((L + R) / 2, T - 10, self.title, font, anchor="cb")
        legend = [e[1:] for e in sorted(legend, key=lambda e: e[0])]
        if self.show_legend and legend:
            line_h, pad, sample = 7 * font + 8, 8, 40
            wbox = sample + 3 * pad + max(len(e[3]) for e in legend) * 6 * font
            hbox = len(legend) * line_h + 2 * pad
            bx, by = R - wbox - 10, T + 10
            cv.rect(bx, by, bx + wbox, by + hbox, (255, 255, 255), 0.85)
            cv.polyline([bx, bx + wbox, bx + wbox, bx, bx], [by, by, by + hbox, by + hbox, by], (180, 180, 180), 1)
            for i, e in enumerate(legend):
                cy = by + pad + i * line_h + line_h // 2
                if e[0] == "line":
                    cv.polyline([bx + pad, bx + pad + sample], [cy, cy], e[1], 2, e[2])
                elif e[0] == "band":
                    cv.rect(bx + pad, cy - 6, bx + pad + sample, cy + 6, e[1], e[2])
                else:
                    cv.marks([bx + pad + sample / 2], [cy], e[1], 5, 1.0)
                cv.text(bx + 2 * pad + sample, cy, e[3], font, anchor="lc")


class Figure:
    """``nrows`` charts stacked in one column (what ``plt.subplots(nrows, 1, figsize=..., sharex=...)`` returns)."""

    def __init__(self, nrows: int = 1, figsize: tuple[float, float] = (12, 5), sharex: bool = False, dpi: int = 150):
        self.axes = [Axes() for _ in range(nrows)]
        self.figsize, self.sharex, self.dpi = figsize, sharex, dpi
        self.title = ""

    def suptitle(self, s: str, **_):
        self.title = s

    def render(self, dpi: int | None = None) -> Canvas:
        dpi = dpi or self.dpi
        W, H = int(round(self.figsize[0] * dpi)), int(round(self.figsize[1] * dpi))
        font = max(1, int(round(dpi / 75)))
        cv = Canvas(W, H)
        title_lines = self.title.split("\n") if self.title else []
        top = 20 + len(title_lines) * (7 * font + 8)
        for i, line in enumerate(title_lines):
            cv.text(W / 2, 14 + i * (7 * font + 8), line, font, anchor="ct")
        left, right, gap = 60 + 6 * font * 9, 30, 28 + (14 * font if not self.sharex else 0)
        n = len(self.axes)
        bottom_margin = 40 + 14 * font
        per_title = [(7 * font + 14) if ax.title else 0 for ax in self.axes]
        avail = H - top - bottom_margin - gap * (n - 1) - sum(per_title)
        h_each = avail / n
        x_range = None
        if self.sharex and n > 1:
            lims = [ax.data_limits()[:2] for ax in self.axes]
            x_range = (min(l[0] for l in lims), max(l[1] for l in lims))
        y = top
        for i, ax in enumerate(self.axes):
            y += per_title[i]
            box = (left, int(y), W - right, int(y + h_each))
            ax.draw(cv, box, x_range, font, x_tick_labels=(not self.sharex) or i == n - 1)
            y += h_each + gap
        return cv

    def savefig(self, path, dpi: int | None = None, **_):
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_bytes(self.render(dpi).png_bytes())


def subplots(nrows: int = 1, ncols: int = 1, figsize: tuple[float, float] = (12, 5), sharex: bool = False):
    """``plt.subplots`` for a single column of charts: returns (figure, axes) — one ``Axes`` for nrows == 1, else a list."""
    if ncols != 1:
        raise ValueError("only a single column of charts is supported")
    fig = Figure(nrows, figsize, sharex)
    return fig, (fig.axes[0] if nrows == 1 else fig.axes)
