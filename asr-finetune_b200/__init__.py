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
from .materialize import iter_parquet, materialize_batch, record_to_arrays, sample_records, to_host_batch, write_parquet
from .sharding import rank_shard, shard_batches

__all__ = ["WhisperFeatureExtractor", "DataCollatorSpeechSeq2SeqWithPadding", "StreamingFrontendCollator",
           "collate_parquet", "labels_fixed_length", "BatchFeature", "slaney_mel_filter_bank", "rank_shard",
           "shard_batches", "materialize_batch", "write_parquet", "iter_parquet", "sample_records", "record_to_arrays",
           "to_host_batch", "_lib"]
__version__ = "0.1.0"
