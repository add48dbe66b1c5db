#!/usr/bin/env python3
"""Record the public surface of the reference's hot-path modules (names, parameters, defaults) as
tests/golden/api_surface.json.  Run in the build container, where /root/reference exists:

    python tests/golden/make_api_surface.py

tests/test_host.py::test_api_surface_matches_the_reference checks python-billiards_b200/billiards against it."""
import inspect
import json
import os
import sys

sys.path.insert(0, "/root/reference/src")
import billiards  # noqa: E402  (the reference)


def params(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        out.append([p.name, p.kind.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


def klass(c):
    names = [n for n in dir(c) if not n.startswith("_")] + ["__init__"]
    return {"methods": {n: params(getattr(c, n)) for n in names if callable(getattr(c, n))},
            "properties": sorted(n for n in names if isinstance(inspect.getattr_static(c, n), property))}


surface = {
    "package": sorted(n for n in ("Billiard", "Disk", "InfiniteWall", "LineSegment", "obstacles", "physics", "simulation")
                      if hasattr(billiards, n)),
    "simulation.Billiard": klass(billiards.simulation.Billiard),
    "obstacles": {c: klass(getattr(billiards.obstacles, c)) for c in ("Obstacle", "Disk", "InfiniteWall", "LineSegment")},
    "physics": {n: params(getattr(billiards.physics, n))
                for n in ("toi_ball_ball", "toi_ball_point", "toi_and_param_ball_segment", "elastic_collision")},
    # instance attributes the reference's callers read (simulation.py:60-112, obstacles.py)
    "Billiard.attributes": ["time", "num", "balls_position", "balls_velocity", "balls_radius", "balls_mass",
                            "balls_initial_time", "balls_initial_position", "obstacles", "toi_table"],
}
b = billiards.Billiard()
surface["Billiard.attributes"] = [a for a in surface["Billiard.attributes"] if hasattr(b, a)]
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "api_surface.json"), "w") as f:
    json.dump(surface, f, indent=1, sort_keys=True)
print(json.dumps(surface, indent=1)[:600])
