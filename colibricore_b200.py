"""`import colibricore_b200 as colibricore`: the reference's Python binding names (PatternModelOptions, UnindexedPatternModel, IndexedPatternModel,
IndexedCorpus, ClassEncoder, ClassDecoder, Pattern) on the B200 library -- see colibri-core_b200/pybinding.py."""
import colibri_core_b200  # noqa: F401  (registers the hyphenated package directory under an importable name)
from colibri_core_b200.pybinding import *  # noqa: F401,F403
from colibri_core_b200.pybinding import (FLEXGRAM, NGRAM, SKIPGRAM, ClassDecoder, ClassEncoder, IndexedCorpus, IndexedPatternModel, Pattern, PatternModelOptions,  # noqa: F401
                                          UnindexedPatternModel)
