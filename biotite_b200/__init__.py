"""
biotite_b200 - B200-native drop-in for the pairwise-alignment hot path of
``biotite.sequence.align`` (``align_optimal`` / ``align_banded``).

The public names mirror the reference's (``src/biotite/sequence/__init__.py``
and ``src/biotite/sequence/align/__init__.py``) for everything on that path.
"""

from .alphabet import Alphabet, AlphabetError, LetterAlphabet
from .sequence import (
    GeneralSequence,
    NucleotideSequence,
    PositionalSequence,
    ProteinSequence,
    PurePositionalSequence,
    Sequence,
)
from .matrix import SubstitutionMatrix
from .alignment import (
    Alignment,
    find_terminal_gaps,
    get_codes,
    get_sequence_identity,
    get_pairwise_sequence_identity,
    get_symbols,
    remove_gaps,
    remove_terminal_gaps,
    score,
)

__version__ = "0.1.0"
from .pairwise import align_optimal, align_ungapped
from .banded import align_banded
from .batch import (
    BatchResult,
    CompositeBatchResult,
    DeviceBatch,
    PackedSequences,
    align_batch_arrays,
    compact_out,
    align_batch_multi,
    align_indexed_batch,
    all_vs_all_scores,
    feng_doolittle_distances,
    align_banded_batch,
    align_optimal_batch,
    pack_sequences,
)
from .localgapped import align_local_gapped, align_local_gapped_batch
from .search import database_shard, search_scores, search_scores_sharded
from .distributed import all_vs_all_scores_sharded
from .statistics import EValueEstimator
from .cigar import CigarOp, read_alignment_from_cigar, write_alignment_to_cigar
from .fasta import alignment_from_fasta, alignment_to_fasta, alignments_to_fasta
