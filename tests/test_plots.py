"""Evaluation plots (eval.py:270-469 of the reference) through the package's own renderer: files, sizes and content."""
import types

import numpy as np
import pytest

from lqr_quadcopter_test_b200 import SuccessCriteria, plots
from lqr_quadcopter_test_b200.eval import Evaluator


def _records(n=600, offset=0.0):
    """Per-step records in the reference's format (eval.py:142-159): a vehicle spiralling in on a moving target."""
    t = 0.01 * np.arange(n)
    target = np.stack([0.5 * t, np.sin(t), 1.0 + 0.0 * t], axis=1)
    quad = target + np.stack([np.exp(-t) * np.cos(3 * t), np.exp(-t) * np.sin(3 * t), -np.exp(-0.5 * t)], axis=1) * (1.0 + offset)
    err = np.linalg.norm(quad - target, axis=1)
    return [{"time": float(t[k]), "step": k, "quadcopter_position": quad[k].tolist(), "target_position": target[k].tolist(),
             "tracking_error": float(err[k]), "on_target": bool(err[k] <= 0.5)} for k in range(n)]


def _evaluator(tmp_path, episodes=2):
    ev = Evaluator(types.SimpleNamespace(name="riccati_lqr"), criteria=SuccessCriteria(), output_dir=tmp_path)
    ev.episode_data_list = [_records(offset=0.5 * i) for i in range(episodes)]
    return ev


def _count(img, rgb, tol=12):
    return int((np.abs(img.astype(int) - np.array(rgb)).max(axis=2) <= tol).sum())


def test_png_round_trip():
    cv = plots.Canvas(40, 30)
    cv.rect(5, 5, 20, 10, (255, 0, 0))
    cv.polyline([0, 39], [29, 0], "tab:blue", 1)
    path = "/tmp/_qt_plot_rt.png"
    open(path, "wb").write(cv.png_bytes())
    img = plots.read_png(path)
    assert img.shape == (30, 40, 3)
    assert (img[7, 10] == (255, 0, 0)).all() and (img[29, 0] == (31, 119, 180)).all() and (img[0, 0] == 255).all()


def test_nice_ticks():
    np.testing.assert_allclose(plots.nice_ticks(0.0, 30.0, 8), [0, 5, 10, 15, 20, 25, 30])
    np.testing.assert_allclose(plots.nice_ticks(-0.12, 0.33, 6), [-0.1, 0.0, 0.1, 0.2, 0.3], atol=1e-12)
    assert len(plots.nice_ticks(1.0, 1.0)) == 1


def test_generate_all_plots_files_and_content(tmp_path):
    ev = _evaluator(tmp_path)
    summary = types.SimpleNamespace(best_episode_idx=0, worst_episode_idx=1)
    saved = ev.generate_all_plots(summary)
    assert [p.name for p in saved] == ["position_tracking_best.png", "tracking_error_best.png",
                                       "position_tracking_worst.png", "tracking_error_worst.png"]
    assert all(p.parent == tmp_path / "plots" and p.stat().st_size > 2000 for p in saved)
    traj = plots.read_png(saved[0])
    assert traj.shape == (1500, 1800, 3)             # 12 x 10 inches at 150 dpi
    # three vehicle curves in their colours, the dashed red target, the green radius band (alpha 0.2 on white)
    for rgb in ((31, 119, 180), (255, 127, 14), (44, 160, 44), (255, 0, 0)):
        assert _count(traj, rgb) > 500, rgb
    assert _count(traj, (204, 230, 204)) > 20000
    err = plots.read_png(saved[1])
    assert err.shape == (750, 1800, 3)               # 12 x 5 inches
    assert _count(err, (0, 0, 255)) > 500 and _count(err, (255, 0, 0)) > 300
    # on-target samples are marked in (blended) green only where the error is inside the radius
    halo = np.abs(err.astype(int) - np.array((127, 191, 127))).max(axis=2) <= 12   # green at alpha 0.5 on white
    assert halo.sum() > 200 and np.nonzero(halo)[1].min() > 500      # the error enters the radius at t = 1.6 s of 6 s
    assert not np.array_equal(plots.read_png(saved[2]), traj)


def test_same_best_and_worst_episode_is_plotted_once(tmp_path):
    ev = _evaluator(tmp_path, episodes=1)
    assert len(ev.generate_all_plots(types.SimpleNamespace(best_episode_idx=0, worst_episode_idx=5))) == 2
    assert len(ev.generate_all_plots()) == 2
    ev.episode_data_list = []
    assert ev.generate_all_plots() == []


def test_plot_interfaces(tmp_path):
    ev = _evaluator(tmp_path)
    fig, axes = ev.plot_trajectory(episode_idx=1, show_radius=False)
    assert len(axes) == 3 and [a.get_ylabel() for a in axes] == ["X Position (m)", "Y Position (m)", "Z Position (m)"]
    assert axes[-1].get_xlabel() == "Time (s)" and "Episode 2" in fig.title and "riccati_lqr" in fig.title
    assert not any(a[0] == "band" for ax in axes for a in ax.artists)
    fig, ax = ev.plot_tracking_error(save_path=tmp_path / "e.png")
    assert ax.get_ylabel() == "Tracking Error (m)" and ax.ylim[0] == 0.0 and (tmp_path / "e.png").exists()
    with pytest.raises(ValueError, match="not found"):
        ev.plot_trajectory(episode_idx=2)
    with pytest.raises(ValueError, match="not found"):
        ev.plot_tracking_error(episode_idx=7)
