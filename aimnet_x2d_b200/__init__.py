"""aimnet_x2d_b200 -- B200 (sm_100a) native hot path of AIMNet-X2D behind the reference's nn.Module API.

Public surface (mirrors ``src/models/__init__.py:9-22`` of the reference):
``GNN``, ``ShellConvolutionLayer``, ``MeanPoolingLayer``, ``MaxPoolingLayer``, ``SumPoolingLayer``,
``MultiHeadAttentionPoolingLayer``, ``create_pooling_layer``, plus the CSR-emitting collation
(``MolBatch``, ``GraphIndex``), the flat-arena trainer pieces (``FlatAdam``) and NCCL helpers.

The CUDA extension ``libax2d.so`` is the only implementation: importing this package loads it and raises
if it is missing (there is no CPU or eager fallback).
"""
from . import _lib

_lib.load()

from .activation import get_activation_function  # noqa: E402
from .collate import GraphIndex, MolBatch, MolData, collate_fn  # noqa: E402
from .gnn import GNN, GNNConfig  # noqa: E402
from .layers import LinearBlock, MultiLayerPerceptron, ShellConvolutionLayer  # noqa: E402
from .losses import WeightedL1Loss, WeightedMSELoss  # noqa: E402
from .inference import (EmbeddingExtractor, GraphedInferenceStep, InferenceStep, OutputFetcher, ShardedOutputWriter,  # noqa: E402
                        merge_rank_outputs)
from .shards import ShardDataset, ShardWriter, shard_batch_indices, write_shard  # noqa: E402
from .optim import FlatAdam  # noqa: E402
from .trainer import GraphedTrainStep, TrainStep  # noqa: E402
from .pooling import (MaxPoolingLayer, MeanPoolingLayer, MultiHeadAttentionPoolingLayer, SumPoolingLayer,  # noqa: E402
                      create_pooling_layer)

__all__ = ["GNN", "GNNConfig", "ShellConvolutionLayer", "LinearBlock", "MultiLayerPerceptron", "MeanPoolingLayer",
           "MaxPoolingLayer", "SumPoolingLayer", "MultiHeadAttentionPoolingLayer", "create_pooling_layer",
           "WeightedL1Loss", "WeightedMSELoss", "GraphIndex", "MolBatch", "MolData", "collate_fn", "FlatAdam",
           "get_activation_function", "InferenceStep", "GraphedInferenceStep", "TrainStep",
           "GraphedTrainStep", "EmbeddingExtractor", "OutputFetcher", "ShardedOutputWriter", "merge_rank_outputs", "ShardDataset",
           "ShardWriter", "shard_batch_indices", "write_shard"]
