"""Import shim: put ``robo-vln_b200/dropin`` (and the repo root) on ``sys.path`` ahead of the
reference checkout and ``robo_vln_baselines.models.seq2seq_highlevel_cma`` /
``...seq2seq_lowlevel`` resolve to the B200 modules (hierarchical_trainer.py:50-51)."""
