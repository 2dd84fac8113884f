#!/usr/bin/env python3
"""Synthetic stand-ins for the reference's downloadable test data (TEST INFRASTRUCTURE).

The reference's data-driven test programs (test/src/scaling.c, asc-bias.c, partial-traversal.c)
read `testdata/2000.tree`, `testdata/2000.fas`, `testdata/246x4465.tree|.fas`, which its test
Makefile fetches from the network (test/Makefile:33-50) - there is none here.  This script writes
files of the same names and shapes with a fixed seed; `make -C oracle dropin` then records what
the REFERENCE library prints for them and tests/test_reference_programs_gpu.py checks that the
same unmodified programs print the same on the device.  (The reference's text fixtures for these
three programs belong to the original files and are not used.)

usage: python oracle/make_testdata.py <output-dir>
"""
import os
import random
import sys


def random_unrooted_tree(rng, names, last_is_tip=True):
    """Newick string of a random binary unrooted tree: (A,B,C) at the top; with last_is_tip the
    third subtree is a single tip (scaling.c:333-335 asserts that root->next->next->back is one)."""
    def blen():
        return f"{rng.uniform(0.005, 0.25):.6f}"

    def join(parts):
        parts = list(parts)
        while len(parts) > 1:
            a = parts.pop(rng.randrange(len(parts)))
            b = parts.pop(rng.randrange(len(parts)))
            parts.append(f"({a}:{blen()},{b}:{blen()})")
        return parts[0]

    names = list(names)
    rng.shuffle(names)
    if last_is_tip:
        third, rest = names[-1], names[:-1]
    else:
        cut = len(names) // 3
        third, rest = join(names[:cut]), names[cut:]
    half = len(rest) // 2
    a, b = join(rest[:half]), join(rest[half:])
    return f"({a}:{blen()},{b}:{blen()},{third}:{blen()});\n"


def random_rooted_tree(rng, names, tip_at_root=False):
    """Strictly binary rooted Newick (reference src/parse_rtree.y:120-160); with tip_at_root one
    child of the root is a single tip (the reference's small.rooted.tip.tree)."""
    def blen():
        return f"{rng.uniform(0.005, 0.25):.6f}"

    def join(parts):
        parts = list(parts)
        while len(parts) > 1:
            a = parts.pop(rng.randrange(len(parts)))
            b = parts.pop(rng.randrange(len(parts)))
            parts.append(f"({a}:{blen()},{b}:{blen()})")
        return parts[0]

    names = list(names)
    rng.shuffle(names)
    if tip_at_root:
        return f"({join(names[:-1])}:{blen()},{names[-1]}:{blen()});\n"
    half = len(names) // 2
    return f"({join(names[:half])}:{blen()},{join(names[half:])}:{blen()});\n"


def alignment(rng, names, sites, alphabet="ACGT", gap="-", p_mut=0.3, p_gap=0.01):
    root = [rng.choice(alphabet) for _ in range(sites)]
    rows = []
    for name in names:
        seq = [(c if rng.random() > p_mut else rng.choice(alphabet)) for c in root]
        seq = [(gap if rng.random() < p_gap else c) for c in seq]
        rows.append((name, "".join(seq)))
    return rows


def write_fasta(path, rows, width=80):
    with open(path, "w") as f:
        for name, seq in rows:
            f.write(f">{name}\n")
            for i in range(0, len(seq), width):
                f.write(seq[i:i + width] + "\n")


def main(out):
    os.makedirs(out, exist_ok=True)
    rng = random.Random(20250117)
    names = [f"taxon{i:04d}" for i in range(2000)]
    open(os.path.join(out, "2000.tree"), "w").write(random_unrooted_tree(rng, names))
    write_fasta(os.path.join(out, "2000.fas"), alignment(rng, names, 300))
    names = [f"seq{i:03d}" for i in range(246)]
    open(os.path.join(out, "246x4465.tree"), "w").write(random_unrooted_tree(rng, names, last_is_tip=False))
    small = [f"sp{i:02d}" for i in range(12)]
    open(os.path.join(out, "small.rooted.tree"), "w").write(random_rooted_tree(rng, small))
    open(os.path.join(out, "small.rooted.tip.tree"), "w").write(random_rooted_tree(rng, small, tip_at_root=True))
    write_fasta(os.path.join(out, "small.fas"), alignment(rng, small, 500, p_mut=0.25))
    rows = alignment(rng, names, 4465, p_mut=0.2)
    write_fasta(os.path.join(out, "246x4465.fas"), rows)
    # the same alignment as interleaved PHYLIP (reference examples/newick-phylip-unrooted)
    with open(os.path.join(out, "246x4465.phy"), "w") as f:
        f.write(f" {len(rows)} {len(rows[0][1])}\n")
        width = 60
        for start in range(0, len(rows[0][1]), width):
            for name, seq in rows:
                f.write((f"{name:<12s}" if start == 0 else "") + seq[start:start + width] + "\n")
            f.write("\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "oracle/_ref/dropin/testdata")
