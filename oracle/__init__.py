"""CPU oracle for the UniIR hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and
there only as the checker / the CPU arm.  The product path (``uniir_b200``) never
imports this package and fails loudly if its CUDA extension is missing.

Parity status (see DESIGN.md "Oracle"):
  * CLIP towers: the arithmetic lives in the un-vendored, unpinned pip dependency
    ``git+https://github.com/openai/CLIP.git`` (reference: src/models/uniir_env.yml:23;
    call sites clip_sf.py:25,26,44,47,66).  It is absent from /root/reference and not
    installable here, and the reference holds no tests or golden vectors -> "parity
    unpinned" by the reference.  We restate the published algorithm (cross-checked
    against the reference's own restatements clip_ff.py:35-59,148-156) and pin the
    restatement against an independent second implementation, ``transformers.CLIPModel``
    (tests/test_oracle_clip.py), and against committed golden vectors in tests/golden/.
  * CLIP_SF loss / fusion / embed API: restated from clip_sf.py:53-168.
  * BLIP_FF / BLIP_SF (oracle/blip_oracle.py): the arithmetic is vendored in the reference (backbone/vit.py,
    med.py, blip_ff.py, blip_sf.py); the restatements are pinned against those UNMODIFIED modules / classes
    imported through shims (tests/test_oracle_blip.py, tests/ref_blip_shim.py): pinned by the reference itself.
  * CLIP_FF (oracle/clip_ff_oracle.py; oracle only, no product path yet): the T5 fusion stack is pinned against
    transformers' own T5Stack, the class the reference instantiates (tests/test_oracle_clip_ff.py).
  * FAISS IndexFlatIP search: faiss-gpu is absent; restated from its documented
    semantics + call sites mbeir_retriever.py:69-103,188-232.
"""
