class MultiLayerPerceptron:  # only used for the ``type(model) is ...`` test in core/sdf/l4casadi.py:235
    pass
