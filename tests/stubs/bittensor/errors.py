class NotVerifiedException(Exception):
    pass
