"""Writes the runnable examples under examples/: case files in the reference's input syntax + Gmsh 2.2 meshes from the synthetic
generators, one per analysis the stand-alone driver covers (analogues of the reference's tutorials ME-TH-AC-001, ME-TH-EL-001 and
ME-ST-EL-002 on the S-cube).  usage: python tools/make_examples.py [cells per face]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multifebe_b200.host import cube_mesh, write_gmsh22, shape

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NAMES = {1: "x0", 2: "xL", 3: "y0", 4: "yL", 5: "z0", 6: "zL"}
BOUNDARIES = "[boundaries]\n6\n" + "".join("%d %d ordinary\n" % (b, b) for b in range(1, 7))
REGION = "[regions]\n1\n\n1 be\n6 1 2 3 4 5 6\nmaterial 1\n0\n"

ROOM = """[problem]
type = mechanics
analysis = harmonic
n = 3D
description = pressure waves in a cubic room of side 3 m (p = 0 at x = 0, p = 1 Pa at x = L, rigid walls elsewhere)

[frequencies]
Hz
lin
12
5.
115.

[settings]
mesh_file_mode = 2 "room.msh"

[materials]
1
1 fluid c 343. rho 1.25

""" + BOUNDARIES + "\n" + REGION + "0\n" + """
[conditions over be boundaries]
boundary 1: 0 (0.,0.)
boundary 2: 0 (1.,0.)
boundary 3: 1 (0.,0.)
boundary 4: 1 (0.,0.)
boundary 5: 1 (0.,0.)
boundary 6: 1 (0.,0.)
"""

COLUMN = """[problem]
n = 3D
type = mechanics
analysis = %(analysis)s
description = unit cube clamped at x = 0, unit normal traction at x = L, sliding lateral faces (1D P-wave column)
%(freq)s
[settings]
mesh_file_mode = 2 "%(mesh)s"

[materials]
1
1 elastic_solid rho 1. mu 1. nu 0.25 xi 0.02

""" + BOUNDARIES + "\n" + REGION + """%(incident)s
[internal points]
3
1 1 0.25 0.5 0.5
2 1 0.50 0.5 0.5
3 1 0.75 0.5 0.5

[export]
real_format = eng_double

[conditions over be boundaries]
boundary 1: 0 %(z)s
            0 %(z)s
            0 %(z)s
boundary 2: 1 %(one)s
            1 %(z)s
            1 %(z)s
boundary 3: 1 %(z)s
            0 %(z)s
            1 %(z)s
boundary 4: 1 %(z)s
            0 %(z)s
            1 %(z)s
boundary 5: 1 %(z)s
            1 %(z)s
            0 %(z)s
boundary 6: 1 %(z)s
            1 %(z)s
            0 %(z)s
"""


def write(name, dat, mesh_name, mesh):
    d = os.path.join(ROOT, "examples", name)
    os.makedirs(d, exist_ok=True)
    write_gmsh22(mesh, os.path.join(d, mesh_name), NAMES)
    open(os.path.join(d, name + ".dat"), "w").write(dat)
    print("examples/%s/%s.dat  (%d nodes, %d elements)" % (name, name, len(mesh.nodes), mesh.n_elem))


write("room", ROOM, "room.msh", cube_mesh(m, shape.QUAD9, L=3.0))
freq = "\n[frequencies]\nrad/s\nlin\n8\n0.5\n7.5\n"
write("column_harmonic", COLUMN % dict(analysis="harmonic", freq=freq, mesh="column.msh", z="(0.,0.)", one="(1.,0.)", incident="0\n"), "column.msh", cube_mesh(m, shape.TRI6))
write("column_static", COLUMN % dict(analysis="static", freq="", mesh="column.msh", z="0.", one="1.", incident=""), "column.msh", cube_mesh(m, shape.QUAD8))

# a soft cubic inclusion in an unbounded elastic medium under a plane P wave: two coupled regions sharing the six faces (the outer region lists them
# reversed), the incident field of the [incident waves] section in the outer region only
INCLUSION = """[problem]
n = 3D
type = mechanics
analysis = harmonic
description = soft cubic inclusion (side 1) in a full space, plane P wave travelling along (sin 30, cos 30, 0) cos 20 + z sin 20

[frequencies]
rad/s
list
3
1.
2.
4.

[settings]
mesh_file_mode = 2 "inclusion.msh"

[materials]
2
1 elastic_solid rho 1. mu 0.25 nu 0.3 xi 0.02
2 elastic_solid rho 1. mu 1. nu 0.25 xi 0.01

""" + BOUNDARIES + """
[bem formulation over boundaries]
""" + "".join("boundary %d: sbie_boundary_mca 0.05\n" % b for b in range(1, 7)) + """
[regions]
2

1 be
6 1 2 3 4 5 6
material 1
0
0

2 be
6 -1 -2 -3 -4 -5 -6
material 2
0
1 1

[incident waves]
1
1
plane
full-space
0 (1.,0.) 0. 0. 0. 30. 20.
0. 0. 0. 0. 0. 0.
elastic p

[export]
real_format = eng_simple
"""
write("inclusion_p_wave", INCLUSION, "inclusion.msh", cube_mesh(m, shape.QUAD9))
