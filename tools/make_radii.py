#!/usr/bin/env python
"""Generate md-structure-factor_b200/radii.txt: ``Z  label  radius_pm`` rows in the format
dens.load_radii parses (reference dens.py:23-35).

The table is built from element data (atomic number, calculated atomic radius in pm; Clementi et
al. values, the same convention the reference's table uses) and expanded to atom-NAME labels,
because compute_sf looks atoms up by name (reference load_traj.py:110).  It is a superset of the
labels in the reference's table (numbered H/C/O/N names, the LLC monomer's named sites, the R3
test particle) plus the GROMACS water names OW/HW*, which the reference's own test systems use
but its table lacks.
"""
import os

ELEMENTS = [  # symbol, Z, radius_pm
    ("H", 1, 53), ("He", 2, 31), ("Li", 3, 167), ("Be", 4, 112), ("B", 5, 87), ("C", 6, 70), ("N", 7, 56),
    ("O", 8, 60), ("F", 9, 42), ("Ne", 10, 38), ("Na", 11, 227), ("Mg", 12, 145), ("Al", 13, 118), ("Si", 14, 111),
    ("P", 15, 98), ("S", 16, 88), ("Cl", 17, 79), ("Ar", 18, 71), ("K", 19, 243), ("Ca", 20, 194), ("Sc", 21, 184),
    ("Ti", 22, 176), ("V", 23, 171), ("Cr", 24, 166), ("Mn", 25, 161), ("Fe", 26, 156), ("Co", 27, 152),
    ("Ni", 28, 149), ("Cu", 29, 145), ("Zn", 30, 142), ("Ga", 31, 136), ("Ge", 32, 125), ("As", 33, 114),
    ("Se", 34, 103), ("Br", 35, 94), ("Kr", 36, 88), ("Rb", 37, 265), ("Sr", 38, 219), ("Y", 39, 212),
    ("Zr", 40, 206), ("Nb", 41, 198), ("Mo", 42, 190), ("Tc", 43, 183), ("Ru", 44, 178), ("Rh", 45, 173),
    ("Pd", 46, 169), ("Ag", 47, 165), ("Cd", 48, 161), ("In", 49, 156), ("Sn", 50, 145), ("Sb", 51, 133),
    ("Te", 52, 123), ("I", 53, 115), ("Xe", 54, 108), ("Cs", 55, 298), ("Ba", 56, 253), ("Pr", 59, 247),
    ("Nd", 60, 206), ("Pm", 61, 205), ("Sm", 62, 238), ("Eu", 63, 231), ("Gd", 64, 233), ("Tb", 65, 225),
    ("Dy", 66, 228),
]
AROMATIC_C = 67      # ring / ester carbons of the LLC monomer carry the smaller radius
ESTER_O = 48


def rows():
    out = [(1, "R3", 90)]                                   # 1-electron test particle of the -RC/-LL modes
    for sym, z, rad in ELEMENTS:
        out.append((z, sym, rad))
    out += [(1, "H%d" % i, 53) for i in range(1, 201)] + [(1, n, 53) for n in ("HSI", "HEST", "HES", "HW", "HW1", "HW2", "HW3")]
    ring = {135, 136, 145}
    out += [(6, "C%d" % i, AROMATIC_C if i in ring else 70) for i in range(1, 201)]
    out += [(6, n, AROMATIC_C) for n in ("CBA", "CB", "CPHE", "CBEN", "CBE", "MEST", "MES", "CSI", "CSI2", "CSI3", "CPH", "CBC", "CBZ")]
    out += [(7, "N%d" % i, 56) for i in range(1, 41)] + [(7, "NZ", 56)]
    out += [(8, "O%d" % i, 60) for i in range(1, 41)] + [(8, n, ESTER_O) for n in ("OEST", "OES", "OS")]
    out += [(8, n, 60) for n in ("OW", "OW1")]
    out += [(11, "NA", 227), (14, "SIL", 111)]
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "md-structure-factor_b200", "radii.txt")
    with open(path, "w") as fh:
        for z, label, rad in rows():
            fh.write("%d\t%s\t%d\n" % (z, label, rad))
    print("wrote", os.path.normpath(path), len(rows()), "rows")
