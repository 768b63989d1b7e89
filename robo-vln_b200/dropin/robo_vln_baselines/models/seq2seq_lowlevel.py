from robovln_b200.seq2seq_lowlevel import Seq2Seq_LowLevel  # noqa: F401
