from robovln_b200.seq2seq_highlevel_cma import Seq2Seq_HighLevel_CMA  # noqa: F401
