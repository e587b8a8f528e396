"""SkillshotGame.get_board (SkillshotGame.py:36-56) is rebuilt on the host by game.render_board from one exported state
row (rendering stays on the host).  tests/golden/boards.npz holds, per tick, the state of the UNMODIFIED reference and the
non-zero cells of the raster its own get_board() drew (oracle/gen_golden.py: boards_scenario): render_board must redraw
every raster cell for cell -- bodies, direction pointers (all the cells floor(-sin * 2.5 + 2.5) can reach, and the case
where the pointer index falls outside the shape and nothing is drawn), projectile crosses, overlaps, board corners."""
import numpy as np

from tests.helpers import load_golden


def dense(cells):
    board = np.zeros((250, 250), dtype=int)
    for x, y, v in cells:
        if v >= 0:
            board[x, y] = v
    return board


def test_render_board_redraws_the_reference_rasters():
    from skillshot_learning_b200.game import render_board
    g = load_golden("boards")
    n, T1 = g["ticks"].shape
    pointers, crosses = set(), 0
    for i in range(n):
        for t in range(T1):
            fields = {k: g[k][i:i + 1, t] for k in ("px", "py", "qx", "qy", "valid", "prot")}
            want = dense(g["board"][i, t])
            got = render_board(fields, 0)
            assert got.dtype == want.dtype and np.array_equal(got, want), (i, t)
            for p in range(2):
                cell = np.argwhere(want[fields["px"][0, p]:fields["px"][0, p] + 5, fields["py"][0, p]:fields["py"][0, p] + 5] == 3 + p)
                pointers.update((p, int(c[0]), int(c[1])) for c in cell)
            crosses += int(fields["valid"].sum())
    # the scenario exercises what it claims to: many pointer cells, drawn projectiles
    assert len(pointers) >= 20 and crosses > 200
