"""omr_a2s_multimodal_transformer_b200 -- B200-native (sm_100a) implementation of the hot path of
mariaalfaroc/omr_a2s_multimodal_transformer: CNN image/spectrogram encoders -> 2-D PE + fusion ->
transformer decoder -> vocabulary projection / cross-entropy -> batched greedy decoding.

The Python classes mirror the reference's ``src/transformer`` surface; every operator runs in
``libomr_b200.so`` (hand-written CUDA, C ABI in ``include/omr_b200.h``).  There is no CPU path.
"""
from .ddp import BucketReducer, DataParallel
from .decoder import Decoder, PositionalEncoding1D
from .encoder import HEIGHT_REDUCTION, WIDTH_REDUCTION, ConvBlock, DepthSepConv2D, DSCBlock, Encoder, MixDropout
from .graph import GraphedTrainStep
from .greedy import BatchedGreedyDecoder, WeightedGreedyDecoder
from .late_fusion import weighted_prediction, weighted_prediction_batch
from .model import (EOS_TOKEN, NUM_CHANNELS, SOS_TOKEN, CrossAttention, MultimodalTransformer, PositionalEncoding2D,
                    Transformer)
from .optim import FusedAdam
from .params import GradArena

__all__ = [
    "Decoder", "PositionalEncoding1D", "Encoder", "ConvBlock", "DSCBlock", "DepthSepConv2D", "MixDropout",
    "PositionalEncoding2D", "CrossAttention", "Transformer", "MultimodalTransformer", "BatchedGreedyDecoder",
    "WeightedGreedyDecoder", "weighted_prediction", "weighted_prediction_batch", "FusedAdam", "GradArena", "GraphedTrainStep", "DataParallel", "BucketReducer", "HEIGHT_REDUCTION", "WIDTH_REDUCTION", "SOS_TOKEN", "EOS_TOKEN", "NUM_CHANNELS",
]
