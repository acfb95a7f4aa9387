"""Stand-in for the third-party `parfor` package the reference's scripts import
(sims/sim_reconciliation.py:25 of the reference; not vendored there, not installed here).

`@parfor(iterable)` runs the decorated function once per element and rebinds the name to the list
of results (reference usage: sims/sim_reconciliation.py:58-96).  The original spreads the elements
over worker processes; here each call already fills the GPU(s) with a batch of frames, so the
elements (SNR points) run one after the other in this process.
"""


def parfor(iterable, *args, **kwargs):
    def decorator(fn):
        return [fn(item) for item in iterable]
    return decorator


pmap = lambda fn, iterable, *a, **k: [fn(item) for item in iterable]  # noqa: E731
