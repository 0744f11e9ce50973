"""movierec: the NeuMF train / ranking-eval hot path of carlamb/MovieRecommender-TF-TRT on B200.

Same module and symbol names as the reference package (`model`, `data_pipeline`, `trainer`,
`util.movielens_utils`); the arithmetic runs in libmovierec_b200.so (hand-written sm_100a CUDA).
"""

__version__ = "1.0"
