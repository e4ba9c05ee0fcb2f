class ANSI(str):
    pass
