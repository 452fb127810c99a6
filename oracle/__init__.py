"""CPU oracle for the simplicial-complex stage.  TEST INFRASTRUCTURE ONLY.

Every module in this package is a plain PyTorch (CPU, fp32 unless a caller asks
for fp64) restatement of one piece of the reference hot path, with the
reference file:line it follows in each docstring.  Nothing under
``topo_audio_autoencoder_b200/`` may import it: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs do, and only as the checker or the timed CPU arm.

Pinning status (see DESIGN.md "Oracle"):

* ``rectifier_oracle``, ``complex_builder_oracle``: PINNED.  Checked bit for bit against the unmodified reference
  ``rectifier.py`` / ``complex_builder.py`` by ``oracle/make_golden.py``; vectors under ``tests/golden/``.
* ``glue_oracle`` (split / active embeddings / penalties), ``gate_oracle.binary_gumbel_train``,
  ``distance_oracle`` (everything after the STFT, including ``compute_distances`` and its two output files),
  ``sccn_oracle`` (the whole forward body of ``GradientSCCNLayer`` / ``GradientSCCN``) and ``decoder_oracle`` (the
  consumer of the SCCN output): PINNED.  ``oracle/make_golden_glue.py`` imports the unmodified reference
  ``encoder.py`` / ``precompute_distances.py`` / ``custom_sccn.py`` / ``decoder.py`` with the stub modules of
  ``oracle/_ref_stubs.py`` for the three absent third-party packages, runs the reference's own code and asserts the
  restatements reproduce it bit for bit; vectors ``tests/golden/ref_*.npz``.
* Still UNPINNED, because the code is not on this machine or does not exist:
  - ``Conv`` arithmetic (TopoModelX, pyt-team/TopoModelX, path-imported, no version pinned): stand-in
    ``neighborhood @ (x @ W)`` on both sides of every comparison;
  - ``MultiScaleSTFT`` (acids-rave): restated with ``torch.stft`` (Hann, hop = scale / 4, centred, reflect, magnitude);
  - ``gate_oracle.hard_concrete``: the reference contains no Hard Concrete code (README prose only); the spec is the
    builder's, after Louizos et al. 2018.
"""
