import logging as _l

warning = _l.warning
info = _l.info
