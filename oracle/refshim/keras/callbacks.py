ModelCheckpoint = object
