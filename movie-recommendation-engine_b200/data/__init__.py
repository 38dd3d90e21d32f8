"""Drop-ins for the parts of the reference's ``data/`` package that sit next to the hot path
(SURVEY.md 8(f) "next" rows): hard-negative sampling on the walk kernel (N3) and the graph builders /
adjacency lists that feed the sampler (N1, N4).  Dataset ETL (CSV loading, feature extraction) stays out
of scope."""
