"""
Build ``biotite_b200/substitution_matrices.json`` from NCBI-format ``.mat`` text
files (the public NCBI/BLAST matrix collection; the reference ships the same
collection under ``src/biotite/sequence/align/matrix_data``).

Only matrices whose row/column symbols are single characters are imported.
Each entry keeps the symbols and the score rows *in file order*; re-ordering
into alphabet order happens at load time in ``SubstitutionMatrix``.

Usage: python tools/import_matrices.py <dir with .mat files>
"""

import json
import sys
from pathlib import Path


def parse(text):
    lines = [ln.strip() for ln in text.splitlines()]
    lines = [ln for ln in lines if ln and not ln.startswith("#")]
    cols = lines[0].split()
    rows = [ln.split()[0] for ln in lines[1:]]
    scores = [[int(v) for v in ln.split()[1:]] for ln in lines[1:]]
    return rows, cols, scores


def main(src):
    out = {}
    for path in sorted(Path(src).glob("*.mat")):
        rows, cols, scores = parse(path.read_text())
        if any(len(s) != 1 for s in rows + cols):
            continue
        if any(len(r) != len(cols) for r in scores):
            continue
        out[path.stem] = {"rows": "".join(rows), "cols": "".join(cols), "scores": scores}
    dst = Path(__file__).resolve().parent.parent / "biotite_b200" / "substitution_matrices.json"
    with open(dst, "w") as f:
        f.write("{\n")
        for k, (name, entry) in enumerate(out.items()):
            f.write(f'"{name}": {json.dumps(entry, separators=(",", ":"))}')
            f.write(",\n" if k + 1 < len(out) else "\n")
        f.write("}\n")
    print(f"{len(out)} matrices -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1])
