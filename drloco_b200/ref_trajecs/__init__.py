"""Reference-motion table compilers (mirror of reference drloco/ref_trajecs/)."""
from .base_ref_trajecs import BaseReferenceTrajectories, MocapTables  # noqa: F401
from .straight_walk_trajecs import StraightWalkingTrajectories  # noqa: F401
from .loco3d_trajecs import Loco3dReferenceTrajectories  # noqa: F401
