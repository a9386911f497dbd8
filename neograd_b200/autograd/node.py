"""Graph node: host-side bookkeeping of the define-by-run graph (reference: neograd/autograd/node.py).
No arithmetic happens here; it exists so that `Tensor.backward()` visits tensors in the same order and with the same
auto-removal rules as the reference (node.py:31-86)."""


class Node:
  __slots__ = ('tens', 'children', 'parents', 'parent_broadcast_shape', 'backward_fn', 'visited', 'lane_done', 'defer_leaves')

  def __init__(self, tens):
    self.tens = tens
    self.children = []                 # nodes that use this tensor as an operand
    self.parents = []                  # operand nodes that produced this tensor
    self.parent_broadcast_shape = None
    self.backward_fn = None
    self.visited = False
    self.defer_leaves = False          # leaf gradients of this op wait until its data gradient has been issued
    self.lane_done = None              # ids of leaf operands whose gradient from this node went to a backward lane

  def add_child(self, other):
    self.children.append(other)

  def add_parent(self, other):
    self.parents.append(other)

  def are_children_visited(self):
    return all(c.visited for c in self.children)

  def are_parents_visited(self):
    return all(p.visited for p in self.parents)

  def visit_all_children(self):
    for c in self.children:
      c.visited = True

  def top_sort(self):
    """Children-before-parents order starting at this node (node.py:31-51), iterative so deep nets cannot hit Python's
    recursion limit.  A node is emitted once all its children are visited; otherwise its unvisited children are
    expanded first -- the same depth-first pre-order the reference's recursion produces."""
    out = []
    stack = [self]
    while stack:
      node = stack.pop()
      if node.visited:
        continue
      if node.are_children_visited():
        node.visited = True
        out.append(node.tens)
        stack.extend(p for p in reversed(node.parents) if not p.visited)
      else:
        # not emitted now: it is reached again through its last child's parent list (as in the reference)
        stack.extend(c for c in reversed(node.children) if not c.visited)
    return out

  def backward(self, retain_graph):
    """node.py:53-86: backward may start at any intermediate node; tensors downstream of it take no part."""
    from .utils import get_graph   # (cheap: once per backward)
    graph = get_graph()
    graph.reset_visited()
    self.visit_all_children()
    order = self.top_sort()
    graph.reset_visited()
    order.pop(0)
    self.visited = True
    self.tens._backward(self, retain_graph, calculate_grads=False)
    for tens in order:
      node = graph.get_node(tens)
      node.visited = True
      tens._backward(node, retain_graph)

  def __repr__(self):
    return f'Node({self.tens})'

  def __str__(self):
    return f'Node( \n{self.tens}\nbackward_fn: {self.backward_fn}\nvisited: {self.visited}\n )'
