"""Drop-in replacement for the reference's `pipeline/ncuts` package (B200 path).

`run_pipeline.py:14-17` does `from ncuts.ncuts_utils import ncuts_chunk, get_merge_pcds` and
`ncuts_utils.py:22` does `from ncuts.normalized_cut import normalized_cut`; both names resolve here
with the reference's signatures.  See INTEGRATION.md.
"""
