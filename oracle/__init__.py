"""CPU oracle for the Qwen3-TTS hot path (TEST INFRASTRUCTURE — never shipped, never measured as product).

What this is
------------
A torch-CPU fp32 / numpy restatement of the hot path of hamptus/mlx-swift-qwen3-tts
(`/root/reference`, Swift on MLX): talker decode + code predictor + sampler + the
speech-tokenizer decoder, written from the Swift sources cited in each docstring
(paths relative to `/root/reference/Sources/Qwen3TTS/`).

PARITY UNPINNED
---------------
The reference cannot be compiled or imported here (no Swift toolchain, no MLX; its
arithmetic lives in the un-vendored `ml-explore/mlx-swift` 0.30.3, rev
4dccaeda1d83cf8697f235d2786c2d72ad4bb925, `Package.resolved:3-11`) and its own tests
(`Tests/Qwen3TTSTests/*`) hold no numeric golden vector, known-answer test or fixture for
this path.  The oracle is therefore pinned only by (a) self-checks against independent
formulations (`torch.nn.functional` convs / SDPA, direct indexing, quantise→dequantise
round trips — see `tests/test_oracle_*.py`), (b) MLX's published semantics restated in
`oracle/mlx_quant.py`, and (c) THIRD-PARTY implementations of the same architectures that ship in
this image's `transformers` and that the oracle reproduces on the same weights
(`tests/test_oracle_*_vs_hf.py`, `tests/test_oracle_speaker_encoder.py`): `Qwen3Model` for the talker
and code-predictor stacks (max error 0.0, prefill and cached decode), Qwen3-Omni's code2wav modules
for SnakeBeta / causal convolutions / ConvNeXt / DecoderResidualUnit / the codec transformer, Mimi for
the RVQ code-to-embedding path (bit for bit) and for the ICL encoder's SEANet CNN, downsampling conv and
nearest-neighbour search (identical code ids), Qwen2.5-Omni's `ECAPA_TimeDelayNet` for the speaker
encoder, and the Metal integration's `_affine_dequantize_tensor` for the MLX packed-weight layout.  Where the REFERENCE deviates from those upstreams (transposed conv trims only its right side;
no sliding window in the codec transformer; zero-padded downsampling conv and bidirectional transformer
in the ICL encoder) the oracle follows the reference and the tests state the deviation.  None of this is
the reference itself: parity claims against the oracle stay "partial" by construction.

Who may import this package: `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs — as the checker or the CPU baseline only.
The product (`mlx-swift-qwen3-tts_b200/`) never imports it and has no CPU fallback.
"""
