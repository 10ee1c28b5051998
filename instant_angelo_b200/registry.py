"""`models.register` / `models.make` of reference models/__init__.py:1-13."""
models = {}


def register(name):
    def decorator(cls):
        models[name] = cls
        return cls
    return decorator


def make(name, config):
    return models[name](config)
