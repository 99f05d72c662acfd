// pluginlib registration is ROS packaging, not arithmetic: no-op here.
#ifndef PLUGINLIB_EXPORT_CLASS
#define PLUGINLIB_EXPORT_CLASS(cls, base) static_assert(sizeof(cls) > 0, "plugin class");
#endif
