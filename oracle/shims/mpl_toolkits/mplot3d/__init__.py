Axes3D = None
