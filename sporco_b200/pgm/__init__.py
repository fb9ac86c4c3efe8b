"""PGM / FISTA solvers running on the B200 engine (mirror of ``sporco.pgm`` for ConvBPDN)."""
