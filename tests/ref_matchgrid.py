"""Plain-Python restatement of StVO::matchGrid (src/matching.cpp:111-258) with real dict-of-lists grids and Python sets,
written independently of the C oracle to cross-check it on small cases (the reference has no tests for this function)."""
import math

INT_MAX = 2 ** 31 - 1


def hamming(a, b):
    return sum(bin(x ^ y).count("1") for x, y in zip(a.tobytes(), b.tobytes()))


def line_cells(x1, y1, x2, y2):   # src/lineIterator.cpp:34-77
    steep = abs(y2 - y1) > abs(x2 - x1)
    if steep:
        x1, y1, x2, y2 = y1, x1, y2, x2
    if x1 > x2:
        x1, x2, y1, y2 = x2, x1, y2, y1
    dx, dy = x2 - x1, abs(y2 - y1)
    error, ystep = dx / 2.0, (1 if y1 < y2 else -1)
    x, y, max_x, out = int(x1), int(y1), int(x2), []
    while not x > max_x:
        out.append((y, x) if steep else (x, y))
        error -= dy
        if error < 0:
            y += ystep
            error += dx
        x += 1
    return out


def grid_get(grid, rows, cols, x, y, w, out):   # src/gridStructure.cpp:65-76
    for x_ in range(max(0, x - w.left), min(cols, x + w.right + 1)):
        for y_ in range(max(0, y - w.up), min(rows, y + w.down + 1)):
            out.update(grid.get((x_, y_), []))


def match_grid(queries, d1, grid, d2, w, ratio, best_lr, rows, cols, dirs2=None, line_sim_th=0.0):
    n1, n2 = len(d1), len(d2)
    m12, m21, dist = [-1] * n1, [-1] * n2, [INT_MAX] * n2
    matches = 0
    for i1 in range(n1):
        best_d, best_d2, best_idx = INT_MAX, INT_MAX, -1
        cand = set()
        if dirs2 is None:
            grid_get(grid, rows, cols, queries[i1][0], queries[i1][1], w, cand)
        else:
            sx, sy, ex, ey = (int(v) for v in queries[i1])
            vx, vy = float(ex - sx), float(ey - sy)
            mag = math.sqrt(vx * vx + vy * vy)
            vx, vy = (vx / mag, vy / mag) if mag != 0 else (float("nan"), float("nan"))
            grid_get(grid, rows, cols, sx, sy, w, cand)
            grid_get(grid, rows, cols, ex, ey, w, cand)
        if not cand:
            continue
        for i2 in cand:
            if dirs2 is not None and abs(vx * dirs2[i2][0] + vy * dirs2[i2][1]) < line_sim_th:
                continue
            d = hamming(d1[i1], d2[i2])
            if best_lr:
                if d < dist[i2]:
                    dist[i2], m21[i2] = d, i1
                else:
                    continue
            if d < best_d:
                best_d2, best_d, best_idx = best_d, d, i2
            elif d < best_d2:
                best_d2 = d
        if best_d < best_d2 * ratio:
            m12[i1] = best_idx
            matches += 1
    if best_lr:
        for i1 in range(n1):
            i2 = m12[i1]
            if i2 >= 0 and m21[i2] != i1:
                m12[i1] = -1
                matches -= 1
    return matches, m12


def points(q_cell, d1, t_cell, d2, w, ratio, best_lr, rows=48, cols=64):
    grid = {}
    for idx, (x, y) in enumerate(t_cell):
        if 0 <= x < cols and 0 <= y < rows:
            grid.setdefault((int(x), int(y)), []).append(idx)
    return match_grid([tuple(int(v) for v in c) for c in q_cell], d1, grid, d2, w, ratio, best_lr, rows, cols)


def lines(q_line, d1, t_line, t_dir, d2, w, ratio, line_sim_th, best_lr, rows=48, cols=64):
    grid = {}
    for idx, (x1, y1, x2, y2) in enumerate(t_line):
        for (x, y) in line_cells(float(x1), float(y1), float(x2), float(y2)):
            if 0 <= x < cols and 0 <= y < rows:
                grid.setdefault((x, y), []).append(idx)
    return match_grid(q_line, d1, grid, d2, w, ratio, best_lr, rows, cols, dirs2=t_dir, line_sim_th=line_sim_th)
