"""Writes tests/golden/matrix_checksums.json: for every NCBI-format matrix file of the reference
(src/biotite/sequence/align/matrix_data/*.mat) the SHA-256 of its ``{(symbol1, symbol2): score}``
dictionary as the reference's ``SubstitutionMatrix.dict_from_str`` defines it (matrix.py:491-501:
row symbols from the left column, column symbols from the top row, the score table transposed).
Run in the authoring container, where /root/reference exists:

    python tests/golden/make_matrix_checksums.py [/root/reference]
"""
import hashlib
import json
import sys
from pathlib import Path


def canonical(entries):
    """One line per symbol pair, sorted: the text that is hashed (also used by the test)."""
    return "\n".join(f"{a}\t{b}\t{score}" for (a, b), score in sorted(entries.items()))


def read_mat(path):
    rows = [line.split() for line in path.read_text().splitlines() if line.strip() and not line.strip().startswith("#")]
    header, body = rows[0], rows[1:]
    left = [r[0] for r in body]
    entries = {}
    for i, a in enumerate(left):
        for j, b in enumerate(header):
            entries[(a, b)] = int(body[j][1 + i])          # transposed: (row i, column j) reads file cell (j, i)
    return entries


if __name__ == "__main__":
    root = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    out = {}
    for path in sorted((root / "src/biotite/sequence/align/matrix_data").glob("*.mat")):
        out[path.stem] = hashlib.sha256(canonical(read_mat(path)).encode()).hexdigest()
    dst = Path(__file__).resolve().parent / "matrix_checksums.json"
    dst.write_text(json.dumps(out, indent=0, sort_keys=True) + "\n")
    print(f"{len(out)} matrices -> {dst}")
