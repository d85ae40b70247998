"""CPU oracle for the Text2Loc coarse cell-retrieval path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this package, and only as the checker / the timed CPU baseline.
Nothing under ``text2loc_b200/`` imports it; the product path is CUDA-only and fails loudly
if its extension is missing.

What is here
  pyg_ops.py       restatement of the four un-vendored PyG ops the reference calls
                   (torch_geometric==1.7.2 / torch-cluster==1.6.0 / torch-scatter==2.0.9:
                   fps, radius, PointConv, global_max_pool) plus the Data/Batch/transforms
                   surface -- their source is NOT under /root/reference, so these follow the
                   published algorithms with every ambiguous choice pinned (see pyg_ops.py).
  stubs.py         installs those + easydict/nltk/matplotlib/numpy-1 shims into sys.modules so
                   the reference's own Python imports unchanged from /root/reference.
  fake_t5.py       deterministic stand-in for the HF tokenizer + T5 encoder (inputs to the path).
  reference_run.py builds the reference's own CellRetrievalNetwork / eval_epoch / run_coarse
                   from /root/reference (this container only; never on the GPU box).
  restate.py       standalone dense restatement of the whole path in torch fp32 / numpy fp64;
                   this is what travels to the GPU box.  It is pinned against outputs of the
                   reference itself (tests/golden/*.npz made by make_golden.py).
  make_golden.py   the committed script that generated tests/golden/.

Parity status: the reference ships no tests or golden vectors, so the pin is
"outputs of the reference's own Python run here, with the PyG ops restated".  The PyG ops
themselves are therefore *parity unpinned* against torch-cluster/torch-scatter binaries
(unavailable offline); every choice they leave open is a named switch in pyg_ops.py.
"""
