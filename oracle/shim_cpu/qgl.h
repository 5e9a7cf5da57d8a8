// No-op stand-in for <qgl.h>: the reference's cpu/src draws with fixed-function GL from the same classes that
// hold the solver (Constraint::draw, Simulation::draw...).  Headless oracle build only; nothing is rendered.
#pragma once
#define GL_POINTS 0
#define GL_LINES 1
#define GL_QUADS 7
#define GL_TRIANGLE_FAN 6
#define GL_FRONT_AND_BACK 0x408
#define GL_LINE 0x1B01
#define GL_FILL 0x1B02
#define GL_BLEND 0xBE2
#define GL_SRC_ALPHA 0x302
#define GL_ONE_MINUS_SRC_ALPHA 0x303
#define GL_PROJECTION 0x1701
#define GL_MODELVIEW 0x1700
#define GL_DEPTH_BUFFER_BIT 0x100
#define GL_COLOR_BUFFER_BIT 0x4000
template <class... A> inline void ps_gl_noop(A...) {}
#define glVertex2f ps_gl_noop
#define glVertex2d ps_gl_noop
#define glBegin ps_gl_noop
#define glEnd ps_gl_noop
#define glColor3f ps_gl_noop
#define glColor4f ps_gl_noop
#define glPushMatrix ps_gl_noop
#define glPopMatrix ps_gl_noop
#define glPointSize ps_gl_noop
#define glTranslatef ps_gl_noop
#define glScalef ps_gl_noop
#define glRotatef ps_gl_noop
#define glPolygonMode ps_gl_noop
#define glMatrixMode ps_gl_noop
#define glLoadIdentity ps_gl_noop
#define glLineWidth ps_gl_noop
#define glViewport ps_gl_noop
#define glOrtho ps_gl_noop
#define glEnable ps_gl_noop
#define glDisable ps_gl_noop
#define glClearColor ps_gl_noop
#define glClear ps_gl_noop
#define glBlendFunc ps_gl_noop
