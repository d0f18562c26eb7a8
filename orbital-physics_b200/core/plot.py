"""Trajectory plotting / video export for a finished engine run (host only, optional).

Consumers of `engine.history` / `engine.objects`, API-compatible in name and
keyword arguments with the reference's core/plot.py (`plot_orbits`,
`render_orbital_mp4`).  matplotlib (and ffmpeg for the video) are optional
dependencies: they are imported on use, so `core.examples` imports without them.
"""
from __future__ import annotations

import numpy as np

_AXES = {"xy": (0, 1), "xz": (0, 2), "yz": (1, 2)}


def _pyplot():
    try:
        import matplotlib
        matplotlib.use("Agg", force=False)
        import matplotlib.pyplot as plt
        return plt
    except Exception as exc:                      # pragma: no cover - depends on the environment
        raise RuntimeError("plotting needs matplotlib, which is not installed") from exc


def trajectories(engine, every_n: int = 1, upto: int | None = None):
    """name -> [T, 3] array of recorded positions (sub-sampled by `every_n`)."""
    out = {}
    for obj in engine.objects:
        pts = engine.history[obj.uuid]
        if upto is not None:
            pts = pts[:upto]
        out[obj.name] = np.asarray(pts, dtype=np.float64).reshape(-1, 3)[:: max(1, int(every_n))]
    return out


def barycenter_track(engine, every_n: int = 1, upto: int | None = None):
    """[T, 3] mass-weighted mean position over the recorded history."""
    tr = trajectories(engine, every_n, upto)
    masses = np.array([float(o.mass) for o in engine.objects])
    T = min(len(v) for v in tr.values()) if tr else 0
    if T == 0:
        return np.empty((0, 3))
    stack = np.stack([tr[o.name][:T] for o in engine.objects])           # [n, T, 3]
    return (masses[:, None, None] * stack).sum(0) / masses.sum()


def _draw(ax, engine, plane, every_n, with_velocity, show_barycenter, barycenter_trail, upto=None):
    i, j = _AXES[plane]
    for obj in engine.objects:
        pts = trajectories(engine, every_n, upto)[obj.name]
        if len(pts) == 0:
            continue
        ax.plot(pts[:, i], pts[:, j], lw=0.8, label=obj.name)
        ax.scatter(pts[-1, i], pts[-1, j], s=12)
        if with_velocity:
            v = np.asarray(obj.velocity, dtype=np.float64)
            ax.annotate("", xy=(pts[-1, i] + v[i], pts[-1, j] + v[j]), xytext=(pts[-1, i], pts[-1, j]),
                        arrowprops=dict(arrowstyle="->", lw=0.6))
    if show_barycenter:
        bc = barycenter_track(engine, every_n, upto)
        if len(bc):
            if barycenter_trail:
                ax.plot(bc[:, i], bc[:, j], "k--", lw=0.6)
            ax.scatter(bc[-1, i], bc[-1, j], c="k", marker="x", s=20, label="barycenter")
    ax.set_xlabel(plane[0] + " [m]")
    ax.set_ylabel(plane[1] + " [m]")
    ax.set_aspect("equal", adjustable="datalim")


def plot_orbits(engine, every_n: int = 1, plane: str = "xy", separate: bool = False, with_velocity: bool = False,
                show_barycenter: bool = False, barycenter_trail: bool = False, save_path: str | None = None):
    """Draw every body's recorded trajectory projected on `plane`."""
    plt = _pyplot()
    if separate:
        n = len(engine.objects)
        fig, axes = plt.subplots(1, max(n, 1), figsize=(4 * max(n, 1), 4), squeeze=False)
        i, j = _AXES[plane]
        for ax, obj in zip(axes[0], engine.objects):
            pts = trajectories(engine, every_n)[obj.name]
            ax.plot(pts[:, i], pts[:, j], lw=0.8)
            ax.set_title(obj.name)
            ax.set_aspect("equal", adjustable="datalim")
    else:
        fig, ax = plt.subplots(figsize=(6, 6))
        _draw(ax, engine, plane, every_n, with_velocity, show_barycenter, barycenter_trail)
        ax.legend(loc="best", fontsize=7)
    if save_path:
        fig.savefig(save_path, dpi=150)
    else:
        plt.show()
    return fig


def render_orbital_mp4(engine, out_path: str = "orbits.mp4", plane: str = "xy", fps: int = 30, duration_s: int = 30,
                       with_velocity: bool = False, show_barycenter: bool = False, barycenter_trail: bool = False,
                       every_n: int = 1):
    """Animate the recorded history into `out_path` (needs matplotlib + ffmpeg)."""
    plt = _pyplot()
    from matplotlib import animation
    T = max((len(engine.history[o.uuid]) for o in engine.objects), default=0)
    frames = max(1, int(fps * duration_s))
    stops = np.unique(np.linspace(1, max(T, 1), frames).astype(int))
    fig, ax = plt.subplots(figsize=(6, 6))

    def frame(k):
        ax.clear()
        _draw(ax, engine, plane, every_n, with_velocity, show_barycenter, barycenter_trail, upto=int(stops[k]))
        return []

    anim = animation.FuncAnimation(fig, frame, frames=len(stops), blit=False)
    anim.save(out_path, writer=animation.FFMpegWriter(fps=fps))
    plt.close(fig)
    return out_path
