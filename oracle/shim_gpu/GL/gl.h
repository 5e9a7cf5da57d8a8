// Headless stand-in for <GL/gl.h> (pulled in by <cuda_gl_interop.h>). Oracle build only.
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef float GLfloat;
typedef int GLsizei;
typedef long GLsizeiptr;
typedef long GLintptr;
