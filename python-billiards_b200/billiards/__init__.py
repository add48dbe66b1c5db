"""billiards (B200 edition): a 2D physics engine for hard disks whose collision hot path runs on an NVIDIA B200.

Drop-in for the public API of markus-ebke/python-billiards (``/root/reference/src/billiards/__init__.py:13-25``):

    import billiards
    bld = billiards.Billiard(obstacles=[billiards.InfiniteWall((0, -1), (0, 1), exterior="right")])
    bld.add_ball((3, 0), (0, 0), radius=0.2, mass=1)
    bld.add_ball((6, 0), (-1, 0), radius=1, mass=100**5)
    bld.evolve(16)      # -> (157080, 157079): 314159 collisions

plus ``BilliardEnsemble`` for many independent billiards sharded over GPUs.  The arithmetic is done by the
hand-written sm_100a kernels in ``libbilliards_b200.so`` (C ABI: ``include/billiards_b200.h``); importing the
package does not need a GPU, using it does -- there is no CPU fallback.
"""

from collections import namedtuple as _namedtuple

from . import obstacles, physics, simulation
from .ensemble import BilliardEnsemble
from .obstacles import Disk, InfiniteWall, LineSegment, Obstacle
from .simulation import Billiard

__version__ = "1.0.0.dev0+b200"
# same shape as the reference's tuple (__init__.py:31-66 there): (major, minor, micro, releaselevel, serial)
__version__info__ = _namedtuple("VersionInfo", ["major", "minor", "micro", "releaselevel", "serial"])(1, 0, 0, "dev", 0)
__all__ = ["obstacles", "physics", "simulation", "Billiard", "BilliardEnsemble", "Disk", "InfiniteWall", "LineSegment",
           "Obstacle"]
