"""CPU oracle for the simplicial-complex stage.  TEST INFRASTRUCTURE ONLY.

Every module in this package is a plain PyTorch (CPU, fp32 unless a caller asks
for fp64) restatement of one piece of the reference hot path, with the
reference file:line it follows in each docstring.  Nothing under
``topo_audio_autoencoder_b200/`` may import it: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs do, and only as the checker or the timed CPU arm.

Pinning status (see DESIGN.md "Oracle"):

* ``rectifier_oracle``, ``complex_builder_oracle``: PINNED.  Checked bit for bit
  against the unmodified reference ``rectifier.py`` / ``complex_builder.py``
  (importable in the authoring container) by ``oracle/make_golden.py``; the
  resulting vectors are committed under ``tests/golden/``.
* ``glue_oracle`` (split / active embeddings / penalties): restated from
  ``encoder.py``; the functions are pure torch but the module cannot be imported
  (unused ``toponetx`` import), so they are pinned only by reading.
* ``gate_oracle.binary_gumbel_train``: restated from ``encoder.py:33-41`` with
  the Gumbel noise injected.  ``gate_oracle.hard_concrete``: PARITY UNPINNED --
  the reference contains no Hard Concrete code (README prose only); the spec is
  the builder's, after Louizos et al. 2018.
* ``sccn_oracle``: PARITY UNPINNED -- the arithmetic of ``Conv`` lives in
  TopoModelX (pyt-team/TopoModelX, path-imported, no version pinned, absent from
  the machine).  ``custom_sccn.py:62-138`` is restated on a stand-in
  ``Conv = neighborhood @ (x @ W)``.
* ``distance_oracle``: the pair reduction (``precompute_distances.py:11-49``) is
  pinned by reading; ``MultiScaleSTFT`` is acids-rave (absent, unpinned) and is
  restated with ``torch.stft``.
"""
