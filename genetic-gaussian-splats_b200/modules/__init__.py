"""Drop-in `modules` package: same module and function names as the reference's
`modules/` so run_ggs.py / run_sags.py import and run unchanged, with the render +
fitness hot path served by libggs_b200.so (sm_100a).  Put this directory's parent
(`genetic-gaussian-splats_b200/`) on sys.path ahead of the reference."""
