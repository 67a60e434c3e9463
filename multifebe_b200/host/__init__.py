from .mesh import Mesh, read_gmsh22, write_gmsh22, cube_mesh, halfspace_patch, two_box_mesh, without_parts, mirror_mesh
from .model import (Model, Material, symmetry_planes, InternalPointsModel, ME_TH_EL_001_BCS, cube_bcs, column_analytic_u,
                    Fluid, FluidModel, room_bcs, room_analytic, Poro, PoroModel)
from . import shape
from .incident import plane_wave, element_incident, plane_wave_fluid, element_incident_fluid
from .multiregion import MultiRegionModel, Region, SOLID, FLUID
