// empty stand-in for <GL/glu.h>; oracle build only
