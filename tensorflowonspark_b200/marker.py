"""Sentinel objects travelling through the feed queues (reference: tensorflowonspark/marker.py:11-18)."""


class Marker(object):
  """Base class of in-band control markers."""

  def __repr__(self):
    return "<{}>".format(type(self).__name__)

  def __eq__(self, other):
    return type(self) is type(other)

  def __hash__(self):
    return hash(type(self).__name__)


class EndPartition(Marker):
  """Pushed after the last row of an inference partition so the consumer can flush a short batch."""


class Rows(Marker):
  """A chunk of rows shipped as ONE queue item (one pickled RPC per chunk instead of per row)."""

  def __init__(self, rows):
    self.rows = rows

  def __eq__(self, other):
    return isinstance(other, Rows) and self.rows == other.rows

  def __hash__(self):
    return hash(("Rows", len(self.rows)))


class RingBlock(Marker):
  """A block of rows that travelled through the shared-memory ring instead of the queue.

  Only the descriptor crosses the manager: ``pos`` is the ring position, ``nrows`` the number of
  rows, ``layout`` the per-column (offset, nbytes, dtype, row_shape) table."""

  def __init__(self, pos, nrows, layout):
    self.pos, self.nrows, self.layout = pos, nrows, layout

  def __eq__(self, other):
    return isinstance(other, RingBlock) and (self.pos, self.nrows) == (other.pos, other.nrows)

  def __hash__(self):
    return hash(("RingBlock", self.pos))
