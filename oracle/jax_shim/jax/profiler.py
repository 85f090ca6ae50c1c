def annotate_function(func=None, name=None, **_kw):
    if func is None:
        return lambda f: f
    return func
