"""
Extract the reference's golden vectors for the align_optimal / align_banded
path into small fixtures that travel with this repo (``/root/reference`` does
not exist on the GPU box).

Sources (read-only, reference test *data*, no reference source code is copied):
  tests/sequence/data/cas9.fasta                               10 Cas9 sequences
  tests/sequence/data/legacy_consistency/params.json           parameters per file
  tests/sequence/data/legacy_consistency/legacy_alignments.tar.gz
        630 FASTA alignments: 360 align_optimal + 180 align_banded + 90 align_local_gapped
        (the X-drop extension)
  README.rst:108-115                                           avidin/streptavidin (C1)

Output: tests/golden/legacy_alignments.json.gz

Usage: python tests/golden/make_golden.py [/root/reference]
"""

import gzip
import io
import json
import re
import sys
import tarfile
from pathlib import Path


def read_fasta(text):
    entries, header, chunks = [], None, []
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            if header is not None:
                entries.append((header, "".join(chunks)))
            header, chunks = line[1:].strip(), []
        else:
            chunks.append(line)
    if header is not None:
        entries.append((header, "".join(chunks)))
    return entries


def main(ref_root):
    ref = Path(ref_root)
    data = ref / "tests" / "sequence" / "data"
    cas9 = read_fasta((data / "cas9.fasta").read_text())
    legacy = data / "legacy_consistency"
    params = json.loads((legacy / "params.json").read_text())
    cases = []
    with tarfile.open(legacy / "legacy_alignments.tar.gz", "r:gz") as tar:
        for name in sorted(params):
            entry = params[name]
            if entry["method"] not in ("align_optimal", "align_banded", "align_local_gapped"):
                continue
            with io.TextIOWrapper(tar.extractfile(name), encoding="utf-8") as f:
                rows = read_fasta(f.read())
            cases.append({
                "file": name,
                "method": entry["method"],
                "seq_indices": entry["seq_indices"],
                "params": entry["params"],
                "gapped": [rows[0][1], rows[1][1]],
            })

    # C1: the alignment printed in README.rst (three 70-column blocks)
    readme = (ref / "README.rst").read_text().splitlines()
    start = next(k for k, ln in enumerate(readme) if ln.strip().startswith("MVHATSPLLL"))
    block = [ln.strip() for ln in readme[start:start + 8] if ln.strip()]
    top = "".join(block[0::2])
    bottom = "".join(block[1::2])
    assert len(top) == len(bottom) and re.fullmatch(r"[A-Z-]+", top + bottom)

    out = {
        "cas9": [{"header": h, "sequence": s} for h, s in cas9],
        "matrix": "BLOSUM62",
        "cases": cases,
        "readme_avidin_streptavidin": {
            "gapped": [top, bottom],
            "params": {"gap_penalty": [-10, -1], "terminal_penalty": False,
                       "local": False, "max_number": 1000},
        },
    }
    dst = Path(__file__).resolve().parent / "legacy_alignments.json.gz"
    with gzip.GzipFile(dst, "wb", mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print(f"{len(cases)} cases -> {dst} ({dst.stat().st_size} bytes)")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
