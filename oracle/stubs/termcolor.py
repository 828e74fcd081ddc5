def colored(text, color=None, *a, **k):
    return str(text)
