"""fn_ssl_b200 -- B200-native (sm_100a) forward hot path of FN-SSL / IPDnet.

Drop-in module surface (same names as the reference, Audio-WestlakeU/FN-SSL):
    fn_ssl_b200.Module            STFT, AddChToBatch, RemoveChFromBatch, forgetting_norm
    fn_ssl_b200.Model             FNblock, FN_SSL, FN_lightning
    fn_ssl_b200.FixedAarryIPDnet  FNblock, CausCnnBlock, IPDnet
    fn_ssl_b200.IPDnet2           OnlineSpatialNet, SpatialNetLayer, FreqInverse, CausalConv1d, Mamba (parameter holder)
plus the fused end-to-end pipelines (fn_ssl_b200.pipeline), the multi-GPU helpers (fn_ssl_b200.distributed) and the
training step (fn_ssl_b200.training: DP-IPD targets, MSE / frame-level PIT losses, LSTM layer / head with backward passes;
FN_SSL / FNblock in train mode).
All arithmetic runs in libfnssl_b200.so (hand-written CUDA, C ABI in include/fnssl_b200.h).
"""
from . import config  # noqa: F401
from .FixedAarryIPDnet import CausalConv1dBlock, CausCnnBlock, FixedArrayIPDnet, IPDnet  # noqa: F401
from .IPDnet2 import IPDnet2_lightning, IPDnet2Pipeline, IPDnet2Stream, OnlineSpatialNet, SpatialNetLayer, data_preprocess_ipdnet2  # noqa: F401
from .Model import FN_SSL, FN_lightning, FNblock, FullNarrowBlock  # noqa: F401
from .Module import (DPIPD, STFT, AddChToBatch, RemoveChFromBatch, SourceDetectLocalize, forgetting_norm,  # noqa: F401
                     pred_ipd_to_doa)
from .pipeline import FNSSLPipeline, IPDnetPipeline, data_preprocess_fnssl, data_preprocess_ipdnet  # noqa: F401
from .streaming import FNSSLStream, IPDnetStream  # noqa: F401
from .training import (FNSSLTrainModule, IPDnetTrainModule, causcnn_train, conv3x3_causal, dpipd_targets, ipd_head_train, ipd_mse_loss,  # noqa: F401
                       ipd_pit_mse_loss, linear_train, lstm_layer)

__version__ = "0.1.0"
