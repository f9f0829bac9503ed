def setup(opt, vocab):
    """Same dynamic factory as the reference's models/__init__.py:1-8 (opt.model names a module holding CapModel)."""
    import importlib
    try:
        mod = importlib.import_module('models.{}'.format(opt.model))
        model = getattr(mod, 'CapModel')(opt, vocab)
    except Exception:
        raise Exception("Model not supported: {}".format(opt.model))
    return model
