def squared_error(a, b):
    return (a - b) ** 2
