"""
Gapped FASTA output / input of alignments: the part of the reference's
``biotite.sequence.io.fasta`` that sits next to the alignment path
(``io/fasta/convert.py:210-262`` ``get_alignment`` / ``set_alignment``; line wrapping as in
``io/fasta/file.py``, 80 characters per line).  The gapped strings of whole batches come from
the GPU (``DeviceBatch.gapped_sequences``).
"""

from __future__ import annotations

from .alignment import Alignment

__all__ = ["alignment_to_fasta", "alignments_to_fasta", "alignment_from_fasta"]


def _entry(header, string, chars_per_line):
    if "\n" in header:
        raise ValueError("Header line contains a line break")
    lines = [">" + header]
    if chars_per_line is None:
        lines.append(string)
    else:
        lines.extend(string[k:k + chars_per_line] for k in range(0, len(string), chars_per_line))
    return "\n".join(lines)


def alignment_to_fasta(alignment, seq_names, chars_per_line=80):
    """FASTA text of one alignment: one entry per gapped sequence (``set_alignment``)."""
    gapped = alignment.get_gapped_sequences()
    if len(gapped) != len(seq_names):
        raise ValueError(f"Alignment has {len(gapped)} sequences, but {len(seq_names)} names were given")
    return "\n".join(_entry(h, s, chars_per_line) for h, s in zip(seq_names, gapped)) + "\n"


def alignments_to_fasta(gapped_pairs, names, chars_per_line=80):
    """
    FASTA text of a batch: `gapped_pairs` as returned by ``DeviceBatch.gapped_sequences``,
    `names` one ``(name1, name2)`` per pair.
    """
    if len(gapped_pairs) != len(names):
        raise ValueError(f"{len(gapped_pairs)} alignments, but {len(names)} name pairs were given")
    out = []
    for (g1, g2), (n1, n2) in zip(gapped_pairs, names):
        out.append(_entry(n1, g1, chars_per_line))
        out.append(_entry(n2, g2, chars_per_line))
    return "\n".join(out) + "\n"


def alignment_from_fasta(text, sequence_factory, additional_gap_chars=("_",)):
    """Parse gapped FASTA text into an :class:`Alignment` (``get_alignment``)."""
    strings, current = [], None
    for line in text.splitlines():
        line = line.strip()
        if not line or line.startswith(";"):
            continue
        if line.startswith(">"):
            current = []
            strings.append(current)
        elif current is not None:
            current.append(line)
    strings = ["".join(parts) for parts in strings]
    for char in additional_gap_chars:
        strings = [s.replace(char, "-") for s in strings]
    return Alignment.from_strings(strings, sequence_factory)
