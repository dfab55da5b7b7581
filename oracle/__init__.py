"""CPU oracle for the translation-transformer inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package
(`translation_transformer_b200/`) may import, call or execute anything in
this directory.  The only allowed users are `tests/`, `__graft_entry__.smoke()`
and the `cpu_baseline` / `--impl reference` legs of `bench.py`, and there only
as the checker (or the timed CPU baseline), never as the product.

The oracle restates, in plain fp32 torch-on-CPU math, the algorithm of

  * `src/model/modules.py`            (VanillaTransformer.encode_src / decode_tgt)
  * `src/model/embeddings.py`         (TokenEmbedding, PositionalEncoding)
  * `src/utils/drafting.py`           (make_drafts)
  * `src/decoding/speculative_decoding.py`  (greedy + beam-search speculative)
  * `src/decoding/standard_decoding.py`     (plain greedy + beam search)

of Academich/translation-transformer.  Parity is PINNED: the golden vectors in
`tests/golden/` were produced by importing the unmodified reference from
`/root/reference` (script: `tests/golden/make_golden.py`) and
`tests/test_oracle_golden.py` checks every oracle function against them.
"""
