#!/usr/bin/env python
"""Stand-in for the `triangle` binary the reference's meshes were made with
(`./triangle -p -q30 -a0.05 mesh5.poly`, last line of resources/mesh5.1.node):

    python scripts/triangle.py -q30 -a0.05 mesh5.poly            # writes mesh5.1.node, mesh5.1.ele, mesh5.1.poly
    python scripts/triangle.py -q30 -a0.002 --box-with-hole 60 out  # the squirmer domain without an input file

Uses fluidsim_b200.meshgen (conforming Delaunay + Ruppert-style refinement).  A `.poly` file with a vertex count
of 0 takes its vertices from the `.node` file of the same stem."""
import argparse, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import fluidsim_b200 as fb


def write_poly(path, segments, markers, holes):
    with open(path, "w") as f:
        f.write("0  2  0  1\n")
        f.write(f"{len(segments)}  1\n")
        for i, ((a, b), m) in enumerate(zip(segments, markers)):
            f.write(f"{i + 1:4d}    {int(a) + 1:4d}  {int(b) + 1:4d}    {int(m)}\n")
        f.write(f"{len(holes)}\n")
        for i, (x, y) in enumerate(holes):
            f.write(f"{i + 1:4d}   {float(x)!r}  {float(y)!r}\n")
        f.write("# written by fluidsim_b200 scripts/triangle.py\n")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-q", dest="min_angle", type=float, default=20.0, help="minimum angle in degrees (Triangle's -q)")
    ap.add_argument("-a", dest="max_area", type=float, default=None, help="maximum triangle area (Triangle's -a)")
    ap.add_argument("-p", action="store_true", help="accepted for compatibility (a .poly file is always a PSLG)")
    ap.add_argument("--box-with-hole", type=int, default=0, metavar="N", help="generate the unit box with an N-gon hole instead of reading a file")
    ap.add_argument("--on-circle", action="store_true", help="move points created on marker-2 segments onto the circle r=0.25")
    ap.add_argument("file", help="input .poly (or the output stem with --box-with-hole)")
    args = ap.parse_args()
    curves = {2: (0.5, 0.5, 0.25)} if args.on_circle else None
    if args.box_with_hole:
        v, vm, s, sm, h = fb.box_with_hole_pslg(args.box_with_hole)
        stem = args.file
    else:
        stem = args.file[:-5] if args.file.endswith(".poly") else args.file
        node = stem + ".node"
        v, vm, s, sm, h = fb.read_poly_full(stem + ".poly", node if os.path.exists(node) else None)
    P, M, T, S, SM = fb.triangulate(v, vm, s, sm, h, min_angle=args.min_angle, max_area=args.max_area, curves=curves)
    out = stem + ".1"
    fb.write_node(out + ".node", P, M)
    fb.write_ele(out + ".ele", T)
    write_poly(out + ".poly", S, SM, h)
    a, b, c = P[T[:, 0]], P[T[:, 1]], P[T[:, 2]]
    area = 0.5 * ((b[:, 0] - a[:, 0]) * (c[:, 1] - a[:, 1]) - (c[:, 0] - a[:, 0]) * (b[:, 1] - a[:, 1]))
    print(f"{out}.node/.ele/.poly: {len(P)} vertices, {len(T)} triangles, {len(S)} segments, largest area {area.max():.3g}")


if __name__ == "__main__":
    main()
