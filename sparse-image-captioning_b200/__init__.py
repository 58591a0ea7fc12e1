"""B200-native ORT / ACORT captioning hot path (drop-in for the masked layers, relation-transformer
encoder/decoder and decode entry points of jiahuei/sparse-image-captioning).

Import as ``sparse_caption_b200`` (see the shim at the repo root).  Submodules mirror the reference files:
``masked_layer``, ``sampler``, ``prune`` (sparse_caption/pruning), ``relation_transformer`` (sparse_caption/models),
plus ``engine`` (inference orchestration), ``kernels`` (tensor wrappers) and ``lib`` (ctypes binding of the C ABI).
"""
__version__ = "0.1.0"
