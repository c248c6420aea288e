"""Multi-GPU commit: column-sharded LDE -> all-to-all -> row-sharded hashing -> gathered tree top (placeholder, see below)."""


def bench_main(args, rank, world, local_rank, dist, bench):
    raise SystemExit("multi-GPU path not built yet")
