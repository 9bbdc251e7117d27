"""Small helpers kept under the reference's names.

``pairwise`` has the call signature of the reference's ``utils.pairwise``
(/root/reference/utils.py:9-13): it yields consecutive pairs of an iterable and
is what turns a track order such as ``['instance', 'neck', 'thorax']`` into the
limb list ``[('instance', 'neck'), ('neck', 'thorax')]``.
"""


def pairwise(iterable):
    """s -> (s0, s1), (s1, s2), (s2, s3), ...  (reference: utils.py:9-13)."""
    it = iter(iterable)
    try:
        prev = next(it)
    except StopIteration:
        return
    for cur in it:
        yield prev, cur
        prev = cur
