"""ADMM solvers running on the B200 engine (mirror of the ``sporco.admm`` package for the
ConvBPDN path)."""
