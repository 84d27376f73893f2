"""B200-native (sm_100a) Whisper log-mel frontend + speech seq2seq padding collator.

Drop-in for the one data-parallel hot path of asr4memory/asr-finetune:
`WhisperFeatureExtractor(...)` as called by finetune/prepare_dataset and the training collators, and
`DataCollatorSpeechSeq2SeqWithPadding`.  All arithmetic runs in hand-written CUDA kernels behind the C ABI in
`include/wfe.h` (`libwfe.so`, built in-tree); importing the compute classes without that library, or calling
them without a CUDA device, raises — there is no CPU fallback.
"""
from . import _lib
from .collator import (DataCollatorSpeechSeq2SeqWithPadding, StreamingFrontendCollator, collate_parquet,
                       labels_fixed_length)
from .feature_extraction import BatchFeature, WhisperFeatureExtractor, slaney_mel_filter_bank
from .sharding import rank_shard, shard_batches

__all__ = ["WhisperFeatureExtractor", "DataCollatorSpeechSeq2SeqWithPadding", "StreamingFrontendCollator",
           "collate_parquet", "labels_fixed_length", "BatchFeature", "slaney_mel_filter_bank", "rank_shard",
           "shard_batches", "_lib"]
__version__ = "0.1.0"
