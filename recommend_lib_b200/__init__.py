"""daisy-b200: B200-native (sm_100a) BPR-MF / funk-SVD training hot path behind Daisy's surface.

Public surface (mirrors the reference):

* ``BPR(user_num, item_num, factor_num)``            -- BPRMFRecommender.py:28-50
* ``BPRMFRecommender(...).fit()/predict()``           -- the epoch loop BPRMFRecommender.py:157-181
* ``metric_eval(model, test_loader, top_k)``          -- util/metrics.py:88-94
* ``SVD`` / ``RSVD`` (ctor kwargs, ``fit``, ``predict``)  -- util/matrix_factorization.pyx:5-167
* ``NCF(..., model='GMF')`` + ``GMFAdam``           -- NCFRecommender.py:28-125, 255-287 (next row)
* ``BPRFM(..., batch_norm=False, drop_prob=[0, 0])`` + ``FMAdagrad``  -- BPRFMRecommender.py:29-80, 191-219 (N3)
* ``BPRFMBN(..., batch_norm=True, drop_prob)`` + ``FMBNAdagrad``     -- the same script at its defaults
  (bprfm_bn.py; ``BPRFM(batch_norm=True)`` builds it)
* ``NeuMF(..., model in ('MLP', 'NeuMF-end'))`` + ``NeuMFAdam``  -- NCFRecommender.py:28-125, 255-287 with the MLP tower
  (the script's default model; ncf_mlp.py -- outside the scope table, kept because it exists)
* ``Item2Vec`` / ``SGNS`` + ``SGNSAdam``                        -- Item2VecRecommender.py:37-97, 266-277 (N4; item2vec.py)
* ``SVDpp`` (ctor kwargs, ``fit``, ``predict``)                -- util/matrix_factorization.pyx:169-288 (N4; svdpp.py)

All compute goes through the C-ABI library ``libdaisy_b200.so`` (``include/daisy_b200.h``);
there is no CPU fallback: using any of the above without the built library or without
a CUDA device raises.
"""
from importlib import import_module

__all__ = ["BPR", "BPRMFRecommender", "metric_eval", "SVD", "RSVD", "MFRecommender",
           "TripleSampler", "lib"]

_LAZY = {
    "BPR": ".bpr", "BPRMFRecommender": ".bpr", "BPRSGD": ".bpr", "BPRAdam": ".bpr",
    "metric_eval": ".metrics", "topk_candidates": ".metrics", "topk_full": ".metrics", "rank_metrics": ".metrics", "final_kpi": ".metrics",
    "SVD": ".mf", "RSVD": ".mf", "MFRecommender": ".mf", "SVDpp": ".svdpp",
    "NCF": ".ncf", "GMFAdam": ".ncf", "BPRFM": ".bprfm", "FMAdagrad": ".bprfm", "BPRFMBN": ".bprfm_bn", "FMBNAdagrad": ".bprfm_bn",
    "NeuMF": ".ncf_mlp", "NeuMFAdam": ".ncf_mlp", "Item2Vec": ".item2vec", "SGNS": ".item2vec", "SGNSAdam": ".item2vec",
    "TripleSampler": ".sampler", "DeviceTripleSampler": ".sampler",
    "ShardedBPR": ".sharded", "PeerShardedBPR": ".sharded",
    "lib": "._lib",
}


def __getattr__(name):
    if name in _LAZY:
        mod = import_module(_LAZY[name], __name__)
        return mod if name == "lib" else getattr(mod, name)
    raise AttributeError(name)
