"""B200-native Stokes-step engine: Python mirror of the reference's functions over
libfluidsim.so (hand-written CUDA for sm_100a, C ABI in include/fluidsim.h).

Import name: ``fluidsim_b200`` (this directory's name has hyphens; the top-level
``fluidsim_b200`` package points its ``__path__`` here).
"""
from ._lib import FluidsimError, device_count, lib as _clib  # noqa: F401  (loads / builds the .so, fails loudly)
from .mesh import (readNode, readEle, readPoly, write_node, write_ele, find_boundary_pairs,  # noqa: F401
                   filter_wall_pairs, index_sets, square_with_hole, refine_mesh)
from .core import (Mesh, CsrMatrix, solve, buildStiffnessMatrix, buildFemSystem, buildLumpedMassMatrix,  # noqa: F401
                   calculate_divergence, calculate_gradiant, PointLocator, mixing_index, mesh_for,
                   build_mass_and_convection,
                   PRECOND_NONE, PRECOND_JACOBI, PRECOND_AMG, PRECOND_AUTO)
from .stokes import StokesSolver, StokesColor, StokesFood, StokesSweep, food_tracer_grid  # noqa: F401
from .partitioned import PartitionedStokes  # noqa: F401
from .hostmesh import node_block_split, sub_mesh, local_index_sets  # noqa: F401
from .poisson import (PoissonProblem, HeatProblem, apply_periodic_bc, apply_dirichlet_rows,  # noqa: F401
                      add_identity_scaled, helmholtz_smooth)
from .meshgen import triangulate, triangulate_poly, box_with_hole_pslg, read_poly_full  # noqa: F401
from .raster import (raster_field, colorize, splat_points, draw_quiver, colormap_lut, write_png, read_png, write_apng,  # noqa: F401
                     read_apng, FrameSink)


def launch_count() -> int:
    """Kernels launched by libfluidsim since load."""
    return int(_clib.fs_launch_count())
