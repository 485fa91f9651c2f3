"""View-parallel sharding: views are independent units (each iteration of the loop
S/SimulateMultiViewDataset.java:567-613 reads only the shared ground truth and its own PSF), so
view v goes to rank v mod G with no data-path collective; the Philox stream id is the view id, so
the result of a view does not depend on G."""


def views_for_rank(n_views, rank, world_size):
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    return list(range(rank, n_views, world_size))
