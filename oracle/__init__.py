"""CPU restatements of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``multimodalbrainsurvival_b200/`` may import this package.  The
only legal importers are ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` (as the checker or
the timed CPU arm, never as the product path).

Pinning status (see DESIGN.md "Oracle"):
  * cox_oracle       - pinned: SURVEY §4 known-answer vector (generated from the
                       reference ``cox_loss``) + tests/golden/cox_*.npz generated
                       by tools/make_golden.py importing /root/reference.
  * aggregate_oracle - pinned: tests/golden/aggregate_*.npz from the reference's
                       ``extract_features`` / ``get_survival_CI`` bodies.
  * resnet_oracle    - pinned: tests/golden/resnet_*.npz from the reference's
                       ``resnet50().forward_extract`` on a seeded input.
  * mlp_oracle       - pinned: tests/golden/mlp_*.npz (reference RNAOnlyModel /
                       BagHistopathologyRNAModel heads, eval mode).
  * cindex_oracle    - PARITY UNPINNED: lifelines is not vendored by the
                       reference and is absent here; Harrell's C is restated
                       from its published definition.
"""
