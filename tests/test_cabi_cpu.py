"""CPU-only checks of the drop-in boundary: the library loads, exports every
symbol the header declares, and the host-side logic behaves like the reference's
wrappers.  No compute call needs a GPU here."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import biotite_b200 as bt
from biotite_b200 import _cabi
from biotite_b200.banded import crop_band

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    header = (ROOT / "include" / "b200align.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(bt_[a-z_0-9]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    lib = ctypes.CDLL(str(_cabi.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in b200align.h but not exported"
    assert sorted(_cabi.EXPORTED_SYMBOLS) == declared


def test_params_struct_matches_header():
    header = (ROOT / "include" / "b200align.h").read_text()
    body = header[header.index("typedef struct BtParams {"):header.index("} BtParams;")]
    fields = re.findall(r"int32_t\s+([a-z_0-9]+);", body)
    assert fields == [f[0] for f in _cabi.BtParams._fields_]
    assert ctypes.sizeof(_cabi.BtParams) == 4 * len(fields)


def test_no_cpu_fallback_without_gpu():
    lib = _cabi.load()
    if lib.bt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    s1, s2 = bt.NucleotideSequence("ACGT"), bt.NucleotideSequence("ACGT")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        bt.align_optimal(s1, s2, bt.SubstitutionMatrix.std_nucleotide_matrix())


def test_argument_errors_match_reference():
    s1, s2 = bt.NucleotideSequence("ACGT"), bt.NucleotideSequence("ACGT")
    m = bt.SubstitutionMatrix.std_nucleotide_matrix()
    with pytest.raises(ValueError, match="Gap penalty must be negative"):
        bt.align_optimal(s1, s2, m, gap_penalty=5)
    with pytest.raises(ValueError, match="Gap penalty must be negative"):
        bt.align_optimal(s1, s2, m, gap_penalty=(-5, 2))
    with pytest.raises(TypeError, match="either integer or tuple"):
        bt.align_optimal(s1, s2, m, gap_penalty=-5.0)
    with pytest.raises(ValueError, match="at least 1"):
        bt.align_optimal(s1, s2, m, max_number=0)
    with pytest.raises(ValueError, match="do not fit the matrix"):
        bt.align_optimal(bt.ProteinSequence("ACDE"), s2, m)
    with pytest.raises(ValueError, match="out of range"):
        bt.align_banded(s1, s2, m, band=(10, 20))
    with pytest.raises(ValueError, match="Gap penalty must be negative"):
        bt.align_banded(s1, s2, m, band=(-1, 1), gap_penalty=1)


def test_crop_band():
    assert crop_band(10, 20, (-100, 100)) == (-9, 19)
    assert crop_band(10, 20, (3, -2)) == (-2, 3)
    with pytest.raises(ValueError):
        crop_band(10, 20, (20, 25))
    with pytest.raises(ValueError):
        crop_band(10, 20, (-30, -10))


def test_substitution_matrix_order_and_alphabets():
    nuc = bt.SubstitutionMatrix.std_nucleotide_matrix()
    # rows/columns are in alphabet order (ACGTRYWSMKHBVDN), not file order (ATGCSWRYKMBVHDN)
    assert nuc.get_score("A", "A") == 5 and nuc.get_score("A", "C") == -4
    assert nuc.get_score("A", "R") == 1 and nuc.get_score("N", "N") == -1
    assert nuc.score_matrix().dtype == np.int32 and nuc.score_matrix().shape == (15, 15)
    assert not nuc.score_matrix().flags.writeable
    assert nuc.get_alphabet1().extends(bt.NucleotideSequence("ACGT").get_alphabet())
    blosum = bt.SubstitutionMatrix.std_protein_matrix()
    assert blosum.get_score("W", "W") == 11 and blosum.get_score("A", "R") == -1
    assert blosum.get_score("*", "*") == 1 and blosum.score_matrix().min() == -4
    assert blosum.is_symmetric()
    assert "BLOSUM62" in bt.SubstitutionMatrix.list_db()
    pos, p1, p2 = blosum.as_positional(bt.ProteinSequence("ACW"), bt.ProteinSequence("WY"))
    assert pos.shape == (3, 2) and pos.score_matrix()[2, 0] == 11


def test_sequence_code_dtype_rule():
    assert bt.ProteinSequence("ACDE").code.dtype == np.uint8
    assert bt.Sequence.dtype(256) == np.uint8 and bt.Sequence.dtype(257) == np.uint16
    seq = bt.GeneralSequence(bt.Alphabet(range(1000)), [5, 999])
    assert seq.code.dtype == np.uint16
    assert bt.NucleotideSequence("acgtn").get_alphabet() == bt.NucleotideSequence.alphabet_amb
    assert str(bt.ProteinSequence(["Ala", "cys", "D"])) == "ACD"


def test_expand_traces_host_only():
    lib = _cabi.load()
    # path end->start: diag, gap-in-seq1, diag ; first cell (1,1)
    ops = np.array([0, 1, 0], dtype=np.uint8)
    trace_len = np.array([3], dtype=np.int64)
    first = np.array([[1, 1]], dtype=np.int32)
    out = np.empty((3, 2), dtype=np.int64)
    _cabi.check(lib.bt_expand_traces(
        1, _cabi.ptr(trace_len), _cabi.ptr(first), _cabi.ptr(ops),
        _cabi.ptr(np.array([0], dtype=np.int64)), None,
        _cabi.ptr(np.array([0, 3], dtype=np.int64)), _cabi.ptr(out)))
    assert out.tolist() == [[0, 0], [-1, 1], [1, 2]]
