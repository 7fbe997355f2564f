"""oracle/ -- CPU restatement of the reference Triple-GAN training hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
legs may import it, and only as the checker / the timed CPU baseline.  The
product path (`tensorflow-implementation-of-triple-gan_b200/tgan`) never imports
it and fails loudly when its CUDA library is missing.

PARITY UNPINNED: the reference (TensorFlow 1.x, unpinned, API bracket 1.12-1.15)
cannot be imported or executed in this image (no TensorFlow, no tf.contrib for
Python 3.12, no network) and ships no tests, golden vectors or known-answer
values for this path.  The oracle is therefore a restatement
of the reference's Python graph code plus TensorFlow's published op semantics,
cross-checked only against (a) a second, loop-level numpy restatement
(`tf_semantics_np.py`), (b) central finite differences for every gradient, and
(c) hand-computed known answers for the TF `SAME`/`conv2d_transpose` geometry.
"""
