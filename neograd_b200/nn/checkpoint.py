"""Checkpoint (reference: neograd/nn/checkpoint.py): sessions + checkpoints.json + one params file per checkpoint.
Pure file bookkeeping; the parameter files are written by `Model.save` (device -> host float64, the reference's
nested-dict format), so checkpoints are interchangeable with the reference's.

New: `Checkpoint(model, dirpath, optimizer=optim)` also writes the optimizer's state
(`Optimizer.state_dict()`: hyper-parameters, iteration count, per-parameter moments) next to every parameter file as
`<hash>.optim`, and `load()` restores it, so a resumed run continues the exact Adam trajectory.  Without the keyword the
class behaves -- and its files look -- exactly like the reference's."""
import datetime as dt
import json
import os
import secrets
from hashlib import sha256


class Checkpoint:
  def __init__(self, model, dirpath, hash_length=16, optimizer=None):
    self.optimizer = optimizer
    self.session = None
    self.dirpath = None
    assert hash_length > 0 and hash_length <= 64, 'Hash length must be between 1 and 64'
    self.hash_length = hash_length
    self._init_files(dirpath)
    self.model = model

  def _json_path(self):
    return f'{self.dirpath}/checkpoints.json'

  def new_session(self):
    """checkpoint.py:35-45."""
    self.session = self._generate_hash()
    os.mkdir(f'{self.dirpath}/{self.session}')
    return self

  def specify_session(self, session):
    """checkpoint.py:47-60."""
    with open(self._json_path(), 'r') as fp:
      if session in json.load(fp):
        self.session = session
      else:
        raise ValueError(f"Invalid session {session} specified!")

  def add(self, **tracked):
    """checkpoint.py:62-81: only builtin, JSON-serialisable values may be tracked; 'datetime' is reserved."""
    forbidden_attrs = ('datetime')
    allowed_types = (int, float, str, dict, list)
    for attr, val in tracked.items():
      if not isinstance(val, allowed_types):
        raise ValueError(f'Only {allowed_types} can be tracked, if using a Tensor, use data attribute and convert it into native python objects or strings!')
      if attr in forbidden_attrs:
        raise ValueError(f'Attribute {attr} must not be present in forbidden_attrs {forbidden_attrs}, as neograd uses them internally!')
    updated, fname_hash = self._update(tracked)
    self._save(updated, fname_hash)

  def _update(self, new_checkpoint):
    new_checkpoint['datetime'] = str(dt.datetime.now())
    return new_checkpoint, self._generate_hash()

  def _save(self, updated_checkpoint, params_fname_hash):
    """checkpoint.py:100-120."""
    with open(self._json_path(), 'r') as fp:
      existing = json.load(fp)
    if existing.get(self.session) is None:
      existing[self.session] = {}
    existing[self.session][params_fname_hash] = updated_checkpoint
    with open(self._json_path(), 'w') as fp:
      json.dump(existing, fp, indent=4)
    self.model.save(f'{self.dirpath}/{self.session}/{params_fname_hash}.hkl')
    if self.optimizer is not None:
      from .model import _pickler
      with open(f'{self.dirpath}/{self.session}/{params_fname_hash}.optim', 'wb') as fp:
        _pickler.dump(self.optimizer.state_dict(), fp)

  def load(self, params_fname, load_params=True):
    """checkpoint.py:122-148."""
    params_fname_hash = params_fname[:-4] if params_fname.endswith('.hkl') else params_fname
    with open(self._json_path()) as fp:
      session = json.load(fp).get(self.session)
    if session is None:
      raise ValueError(f"Invalid session {self.session}")
    if params_fname_hash not in session:
      raise ValueError(f"File {params_fname} not in current session {self.session} directory! Please specify the session using Checkpoint.specify_session")
    if load_params:
      self.model.load(f'{self.dirpath}/{self.session}/{params_fname_hash}.hkl')
      optim_path = f'{self.dirpath}/{self.session}/{params_fname_hash}.optim'
      if self.optimizer is not None and os.path.exists(optim_path):
        from .model import _pickler
        with open(optim_path, 'rb') as fp:
          self.optimizer.load_state_dict(_pickler.load(fp))
    return session[params_fname_hash]

  def _init_files(self, dirpath):
    """checkpoint.py:150-186: create the directory / checkpoints.json; resume the last session if there is one."""
    dirpath = dirpath.rstrip("/'\\")
    os.makedirs(dirpath, exist_ok=True)
    self.dirpath = dirpath
    if not os.path.exists(self._json_path()):
      open(self._json_path(), 'x').close()
    with open(self._json_path()) as fp:
      contents = fp.read()
    sessions = json.loads(contents) if contents.strip() != '' else {}
    if self.session is None:
      if len(sessions) == 0:
        self.new_session()
      else:
        self.session = list(sessions.keys())[-1]
    with open(self._json_path(), 'w') as fp:
      json.dump(sessions, fp, indent=4)

  def _generate_hash(self):
    return sha256(secrets.token_hex(32).encode('utf-8')).hexdigest()[:self.hash_length]
