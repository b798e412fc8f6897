"""Row-sharded embedding table over the GPUs of one box (placeholder until the
all-to-all path lands; see DESIGN.md "multi-GPU")."""


def bench_main(args, rank, local, world):
    raise NotImplementedError("sharded bench path not implemented yet")
