"""models/__init__.py of the reference (:1-8): `setup(opt, vocab)` looks up `models.<opt.model>.CapModel`.  (The reference's
default opt.model = 'RMN' names no module, so its factory always raises: same behaviour here.)"""
import importlib


def setup(opt, vocab):
    try:
        return getattr(importlib.import_module('models.%s' % opt.model), 'CapModel')(opt, vocab)
    except Exception:
        raise Exception("Model not supported: {}".format(opt.model))
