"""CPU oracle for the Whisper log-mel frontend + seq2seq padding collator.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (`asr-finetune_b200/`)
may import this package; only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` do, and there only as the
checker, never as the thing measured or shipped.

Parity status: the reference repo ships no tests or golden vectors for this path
(SURVEY.md §4, §8c).  The oracle is therefore pinned against OUTPUTS OF THE
REFERENCE ITSELF run in the build container: `transformers.WhisperFeatureExtractor`
(the third-party module that holds the arithmetic; reference pins 4.46.3, the
container has 5.5.0) and the reference's own, unmodified
`DataCollatorSpeechSeq2SeqWithPadding` imported from /root/reference.  The
generating script is `tests/golden/make_golden.py`; its fixtures are committed
under `tests/golden/`.
"""
