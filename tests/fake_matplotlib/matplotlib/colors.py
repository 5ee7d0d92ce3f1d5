def to_rgba(c, alpha=None):
    named = {'black': (0.0, 0.0, 0.0), 'white': (1.0, 1.0, 1.0), 'C0': (0.12156862745098039, 0.4666666666666667, 0.7058823529411765)}
    if isinstance(c, str):
        c = named[c]
    c = tuple(float(v) for v in c)
    a = c[3] if len(c) == 4 else 1.0
    if alpha is not None:
        a = float(alpha)
    return c[:3] + (a,)
