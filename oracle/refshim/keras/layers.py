Convolution2D = MaxPooling2D = Dense = Dropout = Flatten = object
